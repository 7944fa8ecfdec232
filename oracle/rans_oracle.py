"""TEST INFRASTRUCTURE -- pure-Python restatement of the reference's entropy-coder boundary, small cases only.

Follows pMCTF/cpp/rans/rans.cpp (encode_with_indexes :76-139, flush :141-168, decode_stream :279-331, the two bit-level
extensions :37-73), the sub-stream container of pMCTF/cpp/py_rans/py_rans.cpp:67-113,129-163 and pmf_to_quantized_cdf of
pMCTF/cpp/ops/ops.cpp:24-82, on the rANS primitives of rygorous/ryg_rans rans64.h @ c9d162d9 (un-vendored in the reference;
published algorithm restated).  Pinned by tests/test_rans.py against the reference's own C++ compiled into oracle/_ref
(oracle/Makefile `ref`).  Only tests/ may import this file."""
import struct

import numpy as np

L = 1 << 31
PREC, RAW_BITS, RAW_MAX = 16, 4, 15
M64 = (1 << 64) - 1


def _round_away(x):   # std::round: halves away from zero (np.round would round them to even)
    x = float(x)
    return float(np.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


def pmf_to_quantized_cdf(pmf, precision=16):
    one = 1 << precision
    cdf = [0] + [int(_round_away(np.float32(np.float32(p) * np.float32(one))) + 0.5) for p in pmf]
    total = sum(cdf)
    cdf = [(one * c) // total for c in cdf]
    for i in range(1, len(cdf)):
        cdf[i] += cdf[i - 1]
    cdf[-1] = one
    n = len(cdf) - 1
    for i in range(n):
        if cdf[i] == cdf[i + 1]:
            best, donor = None, -1
            for j in range(n):
                f = cdf[j + 1] - cdf[j]
                if f > 1 and (best is None or f < best):
                    best, donor = f, j
            assert donor != -1
            if donor < i:
                for j in range(donor + 1, i + 1):
                    cdf[j] -= 1
            else:
                for j in range(i + 1, donor + 1):
                    cdf[j] += 1
    return cdf


def _intervals(symbols, indexes, cdfs, sizes, offsets):
    out = []
    for s, ti in zip(symbols, indexes):
        ti = int(ti)
        if ti < 0:
            continue
        cdf, esc = cdfs[ti], int(sizes[ti]) - 2
        v, raw = int(s) - int(offsets[ti]), 0
        if v < 0:
            raw, v = -2 * v - 1, esc
        elif v >= esc:
            raw, v = 2 * (v - esc), esc
        out.append((int(cdf[v]), int(cdf[v + 1] - cdf[v]), False))
        if v == esc:
            digits = 0
            while (raw >> (digits * RAW_BITS)) != 0:
                digits += 1
            c = digits
            while c >= RAW_MAX:
                out.append((RAW_MAX, 0, True))
                c -= RAW_MAX
            out.append((c, 0, True))
            for j in range(digits):
                out.append(((raw >> (j * RAW_BITS)) & RAW_MAX, 0, True))
    return out


def _code(intervals):
    words, x = [], L
    for start, rng, raw in reversed(intervals):
        if not raw:
            if x >= ((L >> PREC) << 32) * rng:
                words.append(x & 0xFFFFFFFF)
                x >>= 32
            x = ((x // rng) << PREC) + (x % rng) + start
        else:
            if x >= ((L >> 16) << 32) * (1 << (16 - RAW_BITS)):
                words.append(x & 0xFFFFFFFF)
                x >>= 32
            x = ((x << RAW_BITS) | start) & M64
    words.append((x >> 32) & 0xFFFFFFFF)   # flush: the low word ends up at the lower address
    words.append(x & 0xFFFFFFFF)
    return struct.pack("<%dI" % len(words), *reversed(words))


def encode(calls, parts=1):
    """calls: list of (symbols, indexes, cdfs, sizes, offsets) queued before one flush -> the container bytes."""
    queues = [[] for _ in range(parts)]
    for symbols, indexes, cdfs, sizes, offsets in calls:
        n = len(symbols)
        each = n // parts
        for i in range(parts):
            lo, hi = i * each, (i + 1) * each if i < parts - 1 else n
            queues[i] += _intervals(symbols[lo:hi], indexes[lo:hi], cdfs, sizes, offsets)
    streams = [_code(q) for q in queues]
    widest = max([len(s) for s in streams[:-1]], default=0)
    field = 4 if widest > 65535 else 2
    head = bytes([((parts - 1) << 4) + (1 if field == 2 else 0)])
    for s in streams[:-1]:
        head += struct.pack("<H" if field == 2 else "<I", len(s))
    return head + b"".join(streams)


class Decoder:
    def __init__(self, data, parts=1):
        assert (data[0] >> 4) + 1 == parts
        field = 2 if (data[0] & 15) == 1 else 4
        pos, sizes = 1, []
        for _ in range(parts - 1):
            sizes.append(struct.unpack_from("<H" if field == 2 else "<I", data, pos)[0])
            pos += field
        sizes.append(len(data) - pos - sum(sizes))
        self.parts = []
        for n in sizes:
            w = list(struct.unpack_from("<%dI" % (n // 4), data, pos))
            pos += n
            self.parts.append({"w": w, "p": 2, "x": w[0] | (w[1] << 32)})

    @staticmethod
    def _bits(st, n):
        v = st["x"] & ((1 << n) - 1)
        st["x"] >>= n
        if st["x"] < L:
            st["x"] = (st["x"] << 32) | st["w"][st["p"]]
            st["p"] += 1
        return v

    def decode(self, indexes, cdfs, sizes, offsets):
        parts, n = len(self.parts), len(indexes)
        each, out = n // parts, []
        for i, st in enumerate(self.parts):
            lo, hi = i * each, (i + 1) * each if i < parts - 1 else n
            for ti in indexes[lo:hi]:
                ti = int(ti)
                cdf, size = cdfs[ti], int(sizes[ti])
                esc = size - 2
                target = st["x"] & 0xFFFF
                s = 0
                while int(cdf[s + 1]) <= target:
                    s += 1
                start, rng = int(cdf[s]), int(cdf[s + 1] - cdf[s])
                st["x"] = rng * (st["x"] >> PREC) + (st["x"] & 0xFFFF) - start
                if st["x"] < L:
                    st["x"] = (st["x"] << 32) | st["w"][st["p"]]
                    st["p"] += 1
                v = s
                if v == esc:
                    d = self._bits(st, RAW_BITS)
                    digits = d
                    while d == RAW_MAX:
                        d = self._bits(st, RAW_BITS)
                        digits += d
                    raw = 0
                    for j in range(digits):
                        raw |= self._bits(st, RAW_BITS) << (j * RAW_BITS)
                    v = raw >> 1
                    v = -v - 1 if raw & 1 else v + esc
                out.append(v + int(offsets[ti]))
        return np.array(out, dtype=np.int16)
