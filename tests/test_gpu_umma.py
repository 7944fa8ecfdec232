"""tcgen05 kind::i8 conventions (pmctf_umma.cuh) on the real tensor core vs the numpy emulation: descriptor fields,
K-chunk / 8-row-group strides (incl. the overlapping strides the implicit-GEMM convolution relies on), N = 16/32/48,
accumulate flag, TMEM column offsets, and the full digit-split convolution op list.  Bit-exact (integers)."""
import ctypes as C

import numpy as np
import pytest
import torch

import umma_ref as U

pytestmark = pytest.mark.gpu


def run(A, B, ops, n_blocks, block_stride, out_cols, repeat=1):
    import learned_pmctf_b200 as pkg
    nat = pkg._native
    A = np.ascontiguousarray(A, np.int8)
    B = np.ascontiguousarray(B, np.int8)
    pad = lambda x: np.concatenate([x, np.zeros((-x.size) % 16, np.int8)])  # noqa: E731
    A, B = pad(A), pad(B)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    arr = (nat.UmmaOp * len(ops))(*[nat.UmmaOp(**{"a_unsigned": 0, **o}) for o in ops])
    dops = torch.from_numpy(np.frombuffer(bytes(arr), np.uint8).copy()).cuda()
    out = torch.zeros((n_blocks, 128, out_cols), dtype=torch.int32, device="cuda")
    cyc = torch.zeros(n_blocks, dtype=torch.int64, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    nat.check(nat.lib().pmctf_umma_selftest(dA.data_ptr(), A.size, dB.data_ptr(), B.size, dops.data_ptr(), len(ops), n_blocks,
                                            block_stride, out_cols, out.data_ptr(), repeat, cyc.data_ptr(), err.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "umma_selftest")
    torch.cuda.synchronize()
    assert int(err.item()) == 0, f"MMA did not complete (block {int(err.item()) - 1})"
    return out.cpu().numpy().astype(np.int64), cyc.cpu().numpy()


@pytest.mark.parametrize("n", [16, 32, 48])
def test_plain_gemm(n):
    g = np.random.default_rng(n)
    A = g.integers(-128, 128, 2 * 128 * 16, dtype=np.int8)   # chunk-major: [2][128 rows][16]
    B = g.integers(-128, 128, 2 * 48 * 16, dtype=np.int8)
    ops = [dict(a_off=0, a_lbo=128 * 16, a_sbo=128, b_off=0, b_lbo=48 * 16, b_sbo=128, n=n, d_col=0, accumulate=0)]
    got, _ = run(A, B, ops, 1, 0, 48)
    want = U.emulate(A, B, ops, 1, 0, 48)
    assert np.array_equal(got[..., :n], want[..., :n])
    # cross-check the emulation itself against a plain matmul
    a = A.reshape(2, 128, 16).transpose(1, 0, 2).reshape(128, 32).astype(np.int64)
    b = B.reshape(2, 48, 16).transpose(1, 0, 2).reshape(48, 32).astype(np.int64)
    assert np.array_equal(want[0, :, :n], (a @ b.T)[:, :n])


def test_strided_and_overlapping_chunks():
    g = np.random.default_rng(7)
    A = g.integers(-128, 128, 4096 * 16, dtype=np.int8)
    B = g.integers(-128, 128, 4 * 48 * 16, dtype=np.int8)
    base = dict(b_off=0, b_lbo=768, b_sbo=128, n=16, d_col=0, accumulate=0)
    for a_lbo, a_sbo, a_off in ((16, 128, 0), (38 * 16, 128, 48), (16, 128, 38 * 16 + 32), (2048, 256, 16), (0, 128, 0)):
        ops = [dict(base, a_off=a_off, a_lbo=a_lbo, a_sbo=a_sbo)]
        got, _ = run(A, B, ops, 2, 2048, 16)
        assert np.array_equal(got, U.emulate(A, B, ops, 2, 2048, 16)), (a_lbo, a_sbo, a_off)


def test_accumulate_and_column_offsets():
    g = np.random.default_rng(9)
    A = g.integers(-128, 128, 1024 * 16, dtype=np.int8)
    B = g.integers(-128, 128, 2 * 48 * 16, dtype=np.int8)
    mk = lambda a_off, b_off, n, col, acc: dict(a_off=a_off, a_lbo=16, a_sbo=128, b_off=b_off, b_lbo=768, b_sbo=128, n=n,  # noqa: E731
                                                d_col=col, accumulate=acc)
    ops = [mk(0, 0, 48, 0, 0), mk(160, 0, 32, 16, 1), mk(160, 512, 16, 48, 0), mk(320, 256, 16, 48, 1), mk(64, 256, 16, 64, 0),
           mk(32, 0, 48, 32, 1)]  # every column group is initialised (accumulate=0) before it is accumulated into
    got, _ = run(A, B, ops, 1, 0, 80)
    assert np.array_equal(got, U.emulate(A, B, ops, 1, 0, 80))


def test_mixed_sign_u8_times_s8():
    """kind::i8 with an unsigned A operand and a signed B operand (instruction descriptor a_format = 0, b_format = 1)."""
    g = np.random.default_rng(21)
    A = g.integers(-128, 128, 2 * 128 * 16, dtype=np.int8)
    B = g.integers(-128, 128, 2 * 48 * 16, dtype=np.int8)
    for au in (0, 1):
        ops = [dict(a_off=0, a_lbo=128 * 16, a_sbo=128, b_off=0, b_lbo=48 * 16, b_sbo=128, n=48, d_col=0, accumulate=0, a_unsigned=au)]
        got, _ = run(A, B, ops, 1, 0, 48)
        assert np.array_equal(got, U.emulate(A, B, ops, 1, 0, 48)), f"a_unsigned={au}"
    pitch, n_blocks = 38, 3
    npix = n_blocks * 128 + 2 * pitch + 2 + 8
    Aint = g.integers(-2 ** 22, 2 ** 22 + 1, (npix, 16))
    Wint = g.integers(-2 ** 22, 2 ** 22 + 1, (16, 16, 3, 3))
    Ab = np.concatenate([d.reshape(-1) for d in U.split_digits_twos(Aint)])
    got, _ = run(Ab, U.pack_weights(Wint), U.conv_ops(pitch, npix * 16, twos=True), n_blocks, 2048, 80)
    assert np.array_equal(U.combine_orders(got).reshape(-1, 16), U.conv_exact(Aint, Wint, pitch, n_blocks * 128))


def test_digit_split_convolution_and_timing():
    g = np.random.default_rng(0)
    pitch, n_blocks = 38, 11
    npix = n_blocks * 128 + 2 * pitch + 2 + 8
    Aint = g.integers(-2 ** 22, 2 ** 22 + 1, (npix, 16))
    Wint = g.integers(-2 ** 22, 2 ** 22 + 1, (16, 16, 3, 3))
    plane = npix * 16
    A = np.concatenate([d.reshape(-1) for d in U.split_digits(Aint)])
    B = U.pack_weights(Wint)
    ops = U.conv_ops(pitch, plane)
    got, cyc1 = run(A, B, ops, n_blocks, 2048, 80)
    assert np.array_equal(U.combine_orders(got).reshape(-1, 16), U.conv_exact(Aint, Wint, pitch, n_blocks * 128))
    _, cyc = run(A, B, ops, n_blocks, 2048, 80, repeat=64)
    per_mma = float(np.median(cyc)) / (64 * len(ops))
    print(f"\ntcgen05 kind::i8 M=128 conv op list: {np.median(cyc1):.0f} cycles per 15-MMA block (issue+commit+wait), "
          f"{per_mma:.1f} cycles per MMA steady state (N mix 48/32/16)")
    for n in (16, 32, 48, 96):
        o = [dict(a_off=32 * i, a_lbo=16, a_sbo=128, b_off=0, b_lbo=768 if n <= 48 else 1536, b_sbo=128, n=n, d_col=0, accumulate=1)
             for i in range(48)]
        _, c = run(A, B if n <= 48 else np.concatenate([B, B]), o, 4, 2048, 96, repeat=16)
        print(f"  N={n}: {float(np.median(c)) / (16 * 48):.1f} cycles per MMA (48 back-to-back MMAs x 16)")
