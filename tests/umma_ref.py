"""numpy emulation of the tcgen05 kind::i8 op lists executed by pmctf_umma_selftest, and the builder of the op list
the lifting convolution uses (3 byte digits per operand, 4 tap pairs + tap 8 as a digit pair, 5 accumulator groups).
Test infrastructure."""
import numpy as np

TAP_PAIRS = [((0, 0), (0, 1)), ((1, 0), (1, 1)), ((2, 0), (2, 1)), ((0, 2), (1, 2)), ((2, 2), None)]


def emulate(A, B, ops, n_blocks, block_stride, out_cols):
    """A, B: int8 byte images; ops: dicts with the pmctf_umma_op_t fields.  -> int64 [n_blocks, 128, out_cols]."""
    A_s = np.asarray(A, np.int8).astype(np.int64)
    A_u = np.asarray(A, np.int8).view(np.uint8).astype(np.int64)
    B = np.asarray(B, np.int8).astype(np.int64)
    out = np.zeros((n_blocks, 128, out_cols), np.int64)
    r = np.arange(128)
    for blk in range(n_blocks):
        for op in ops:
            n = op["n"]
            arow = op["a_off"] + blk * block_stride + (r % 8) * 16 + (r // 8) * op["a_sbo"]
            rn = np.arange(n)
            brow = op["b_off"] + (rn % 8) * 16 + (rn // 8) * op["b_sbo"]
            acc = np.zeros((128, n), np.int64)
            A = A_u if op.get("a_unsigned") else A_s
            for c in range(2):
                a = A[(arow + c * op["a_lbo"])[:, None] + np.arange(16)[None, :]]
                b = B[(brow + c * op["b_lbo"])[:, None] + np.arange(16)[None, :]]
                acc += a @ b.T
            sl = slice(op["d_col"], op["d_col"] + n)
            out[blk, :, sl] = acc + (out[blk, :, sl] if op["accumulate"] else 0)
    return out


def split_digits(v):
    """int (|v| <= 2^22) -> three signed-byte digits d0, d1, d2 with v = d0*65536 + d1*256 + d2."""
    v = np.asarray(v, np.int64)
    d2 = ((v + 128) & 0xFF) - 128
    v1 = (v - d2) >> 8
    d1 = ((v1 + 128) & 0xFF) - 128
    d0 = (v1 - d1) >> 8
    assert np.all(np.abs(d0) <= 127) and np.all(d0 * 65536 + d1 * 256 + d2 == v)
    return d0.astype(np.int8), d1.astype(np.int8), d2.astype(np.int8)


def split_digits_twos(v):
    """int (|v| <= 2^22) -> two's complement byte digits: d0 = v >> 16 (signed), u1, u2 in [0, 255] with
    v = d0*65536 + u1*256 + u2 -- what the kernel stores for the activations (plane 0 signed, planes 1, 2 unsigned)."""
    v = np.asarray(v, np.int64)
    u2, u1, d0 = v & 0xFF, (v >> 8) & 0xFF, v >> 16
    assert np.all(np.abs(d0) <= 127) and np.all(d0 * 65536 + u1 * 256 + u2 == v)
    return d0.astype(np.int8), u1.astype(np.uint8).view(np.int8), u2.astype(np.uint8).view(np.int8)


def pack_weights(W):
    """W int [16 co, 16 ci, 3, 3] (|W| <= 2^22) -> int8 image (10240 B, 9728 used):
      * per tap pair tp = 0..3 a [2 chunks][48 rows = (digit, co)][16 ci] block at tp*1536;
      * tap 8 (ky = kx = 2) for the most significant activation digit: the same block format at 6144, second chunk zero;
      * tap 8 for the activation digit PAIR (d1 | d2): [2 chunks][64 rows = (group - 1, co)][16 ci] at 7680: chunk 0
        (times d1) holds weight digit j in row block j, chunk 1 (times d2) holds weight digit j in row block j + 1."""
    d = split_digits(W)
    out = np.zeros((5, 2, 48, 16), np.int8)
    for tp, pair in enumerate(TAP_PAIRS):
        for c, tap in enumerate(pair):
            if tap is None:
                continue
            for j in range(3):
                out[tp, c, j * 16:(j + 1) * 16, :] = d[j][:, :, tap[0], tap[1]]
    last = np.zeros((2, 64, 16), np.int8)
    for j in range(3):
        last[0, j * 16:(j + 1) * 16, :] = d[j][:, :, 2, 2]
        last[1, (j + 1) * 16:(j + 2) * 16, :] = d[j][:, :, 2, 2]
    return np.concatenate([out.reshape(-1), last.reshape(-1), np.zeros(512, np.int8)])


def conv_ops(pitch, plane_bytes, twos=False):
    """Op list of one 128-pixel block of the 16->16 3x3 convolution (14 MMAs): D columns [16*i, 16*i+16) accumulate the
    digit products of order i (weight 2^(32-8i)): a_d x [w0; w1; w2] lands in groups d, d+1, d+2.  Taps 0..7 go pairwise
    (K = 2 taps x 16 channels) per activation digit; tap 8 goes once for digit 0 (second K half against zero weights) and
    once for the digit pair (d1 | d2) (K = 2 digits x 16 channels, N = 64 -> groups 1..4).  The first MMA does not
    accumulate and initialises groups 0..2; groups 3, 4 are zeroed by whoever drained the accumulators before."""
    ops = []
    for tp, (t0, t1) in enumerate(TAP_PAIRS[:4]):
        a_off = (t0[0] * pitch + t0[1]) * 16
        lbo = ((t1[0] * pitch + t1[1]) - (t0[0] * pitch + t0[1])) * 16
        for d in range(3):
            ops.append(dict(a_off=d * plane_bytes + a_off, a_lbo=lbo, a_sbo=128, b_off=tp * 1536, b_lbo=768, b_sbo=128,
                            n=48, d_col=16 * d, accumulate=int(not (tp == 0 and d == 0)), a_unsigned=int(twos and d > 0)))
    t8 = (2 * pitch + 2) * 16
    ops.append(dict(a_off=t8, a_lbo=16, a_sbo=128, b_off=6144, b_lbo=768, b_sbo=128, n=48, d_col=0, accumulate=1, a_unsigned=0))
    ops.append(dict(a_off=plane_bytes + t8, a_lbo=plane_bytes, a_sbo=128, b_off=7680, b_lbo=1024, b_sbo=128, n=64, d_col=16,
                    accumulate=1, a_unsigned=int(twos)))
    return ops


def conv_exact(Aint, Wint, pitch, n_out):
    """Direct integer convolution on the linearised pixel array: out[m, co] = sum_{ci,ky,kx} A[m + ky*pitch + kx, ci] W[co,ci,ky,kx]."""
    Aint = np.asarray(Aint, np.int64)
    Wint = np.asarray(Wint, np.int64)
    out = np.zeros((n_out, 16), np.int64)
    for ky in range(3):
        for kx in range(3):
            out += Aint[ky * pitch + kx: ky * pitch + kx + n_out] @ Wint[:, :, ky, kx].T
    return out


def combine_orders(o):
    """[.., 80] accumulator groups -> exact integer sum  o0*2^32 + o1*2^24 + o2*2^16 + o3*2^8 + o4."""
    o = np.asarray(o, np.int64).reshape(o.shape[:-1] + (5, 16))
    return sum(o[..., i, :] << (32 - 8 * i) for i in range(5))
