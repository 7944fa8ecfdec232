"""CPU check of the digit-split implicit-GEMM formulation itself: the op list the tensor-core convolution issues,
emulated in numpy, reproduces the exact integer 3x3 convolution (no GPU needed)."""
import numpy as np

import umma_ref as U


def test_digit_split_roundtrip():
    v = np.concatenate([np.arange(-2 ** 22, 2 ** 22 + 1, 977), [0, 1, -1, 127, 128, -128, -129, 2 ** 22, -2 ** 22, 32767, 32768]])
    d0, d1, d2 = U.split_digits(v)
    assert np.array_equal(d0.astype(np.int64) * 65536 + d1.astype(np.int64) * 256 + d2, v)
    assert np.abs(d0).max() <= 64


def test_conv_oplist_is_exact():
    g = np.random.default_rng(0)
    pitch, n_blocks = 38, 3
    npix = n_blocks * 128 + 2 * pitch + 2 + 8
    Aint = g.integers(-2 ** 22, 2 ** 22 + 1, (npix, 16))
    Wint = g.integers(-2 ** 22, 2 ** 22 + 1, (16, 16, 3, 3))
    Wint[0, 0, 0, 0], Aint[5, 3] = 2 ** 22, -2 ** 22
    plane = npix * 16
    A = np.concatenate([d.reshape(-1) for d in U.split_digits(Aint)])
    B = U.pack_weights(Wint)
    ops = U.conv_ops(pitch, plane)
    assert len(ops) == 14
    out = U.emulate(A, B, ops, n_blocks, 128 * 16, 80)
    assert np.abs(out).max() < 2 ** 31, "accumulator groups must fit int32"
    got = U.combine_orders(out).reshape(n_blocks * 128, 16)
    assert np.array_equal(got, U.conv_exact(Aint, Wint, pitch, n_blocks * 128))


def test_conv_oplist_twos_complement_activation_digits():
    """Activation digits as plain two's complement bytes (plane 0 signed, planes 1 and 2 unsigned: u8 x s8 MMAs)."""
    g = np.random.default_rng(1)
    pitch, n_blocks = 38, 2
    npix = n_blocks * 128 + 2 * pitch + 2 + 8
    Aint = g.integers(-2 ** 22, 2 ** 22 + 1, (npix, 16))
    Wint = g.integers(-2 ** 22, 2 ** 22 + 1, (16, 16, 3, 3))
    Aint[3, 5], Aint[4, 6] = 2 ** 22, -2 ** 22
    A = np.concatenate([d.reshape(-1) for d in U.split_digits_twos(Aint)])
    out = U.emulate(A, U.pack_weights(Wint), U.conv_ops(pitch, npix * 16, twos=True), n_blocks, 128 * 16, 80)
    assert np.abs(out).max() < 2 ** 31
    assert np.array_equal(U.combine_orders(out).reshape(n_blocks * 128, 16), U.conv_exact(Aint, Wint, pitch, n_blocks * 128))
