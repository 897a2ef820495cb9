"""CPU check of the error bound the tcgen05 filter relies on (pda_eval_tc.cu): the accumulator
v_j = sum_k bf16(u_k) bf16(w_jk) + (x_hi + x_mid + x_lo), accumulated in fp32 sixteen products at a time, stays within
E = cA |u| |w_j| + cB (|u| |w_j| + |x_j|) of the exact value -- emulated in numpy with round-to-nearest AND with
truncating fp32 accumulation (the tensor core's internal rounding is not documented), on adversarial inputs.
The GPU test test_tensor_accumulators_within_error_bound checks the same inequality on the real hardware."""
import zlib

import numpy as np
import pytest


def bf16_rn(x):
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((b + 0x7FFF + ((b >> 16) & 1)) >> 16) << 16
    return (r & 0xFFFFFFFF).astype(np.uint32).view(np.float32)


def split3(x):
    x = np.asarray(x, dtype=np.float32)
    hi = bf16_rn(x)
    r1 = (x - hi).astype(np.float32)
    mid = bf16_rn(r1)
    lo = bf16_rn((r1 - mid).astype(np.float32))
    return hi, mid, lo


def f32_chop(x64):
    """fp64 -> fp32 rounding toward zero"""
    y = x64.astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x64)
    return np.where(over, np.nextafter(y, np.float32(0)), y).astype(np.float32)


def accumulate(U, W, X3, chop):
    """rows of U against rows of W: fp32 accumulator updated once per 16 products (one UMMA K step) + the extra K block"""
    rnd = f32_chop if chop else (lambda z: z.astype(np.float32))
    Ub, Wb = bf16_rn(U).astype(np.float64), bf16_rn(W).astype(np.float64)
    acc = np.zeros((U.shape[0], W.shape[0]), dtype=np.float32)
    for k0 in range(0, U.shape[1], 16):
        acc = rnd(acc.astype(np.float64) + Ub[:, k0:k0 + 16] @ Wb[:, k0:k0 + 16].T)       # products of bf16 are exact in fp32
    if X3 is not None:
        acc = rnd(acc.astype(np.float64) + sum(p.astype(np.float64) for p in X3)[None, :])
    return acc


def coefficients(d):
    cA = 1.02 / 256.0 + d / 2097152.0
    cB = (d // 16 + 5) / 524288.0
    return cA + cB, cB


CASES = {
    "gaussian": lambda rng, n, d: rng.normal(0, 1, (n, d)),
    "all_positive": lambda rng, n, d: rng.random((n, d)) + 0.5,                       # rounding errors do not cancel
    "midpoints": lambda rng, n, d: (1.0 + (rng.integers(0, 128, (n, d)) + 0.5) / 128.0 - 2.0 ** -17)     # just below bf16 ties
                                   * rng.choice([-1.0, 1.0], (n, d)) * 2.0 ** rng.integers(-3, 3, (n, d)),
    "tiny": lambda rng, n, d: rng.normal(0, 3e-3, (n, d)),                            # Xavier-sized tables
    "huge": lambda rng, n, d: rng.normal(0, 40.0, (n, d)),
}


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("kind", ["main_branch", "condition", "bias"])
@pytest.mark.parametrize("chop", [False, True])
def test_filter_bound_holds(d, case, kind, chop):
    rng = np.random.default_rng(zlib.crc32(f"{d}/{case}/{kind}".encode()))
    U = CASES[case](rng, 48, d).astype(np.float32)
    I = CASES[case](rng, 400, d).astype(np.float32)
    pop = (rng.random(400) ** 3).astype(np.float32)
    pop[:20] = 0.0
    bias = rng.normal(0, 0.3, 400).astype(np.float32)
    cAB, cB = coefficients(d)
    S = U.astype(np.float64) @ I.astype(np.float64).T
    if kind == "condition":
        W = (I * pop[:, None]).astype(np.float32)          # fl(pop_j * i_jk), as the conversion kernel forms it
        X3, xa = split3(pop), np.abs(pop.astype(np.float64))
        want = (S + 1.0) * pop.astype(np.float64)[None, :]
    elif kind == "bias":
        W, X3, xa = I, split3(bias), np.abs(bias.astype(np.float64))
        want = S + bias.astype(np.float64)[None, :]
    else:
        W, X3, xa = I, None, np.zeros(400)
        want = S
    v = accumulate(U, W, X3, chop).astype(np.float64)
    un = np.linalg.norm(U.astype(np.float64), axis=1)[:, None]
    wn = np.linalg.norm(W.astype(np.float64), axis=1)[None, :]
    E = cAB * un * wn + cB * xa[None, :] + 1e-30
    ratio = (np.abs(v - want) / E).max()
    assert ratio <= 1.0, ratio


def test_three_bf16_pieces_carry_24_bits():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.random(200000), rng.random(200000) ** 6, rng.normal(0, 0.3, 200000), [0.0, 1.0, -1.0]]).astype(np.float32)
    hi, mid, lo = split3(x)
    err = np.abs(hi.astype(np.float64) + mid.astype(np.float64) + lo.astype(np.float64) - x.astype(np.float64))
    normal = np.abs(x) >= 1e-30
    assert np.all(err[normal] == 0.0)                       # exact for normal-range values
    assert err.max() <= 1e-38                               # subnormal pieces: below the absolute slack of the bound


def test_sign_mirrored_quotient_refinement_is_symmetric():
    """lazy_zero_grad_step4_fast refines R = -1/b instead of r = 1/b: every fma appears with both signs flipped, which
    round-to-nearest mirrors exactly (checked here in fp32 arithmetic through fp64 fma emulation)."""
    rng = np.random.default_rng(2)
    a = (rng.normal(0, 1, 100000) * 10.0 ** rng.integers(-12, 3, 100000)).astype(np.float32)
    b = (rng.random(100000) * 10.0 ** rng.integers(-8, 2, 100000) + 1e-8).astype(np.float32)

    def fma(x, y, z):       # exact product in fp64 (24 x 24 bits), one rounding of the sum to fp32: the cases where the
        return (x.astype(np.float64) * y.astype(np.float64) + z.astype(np.float64)).astype(np.float32)   # fp64 sum rounds twice are sign-symmetric too

    r0 = (1.0 / b.astype(np.float64)).astype(np.float32)
    r = fma(r0, fma(-b, r0, np.float32(1.0)), r0)
    q0 = fma(a, r, np.float32(0.0))
    q = fma(r, fma(-b, q0, a), q0)
    R0 = -r0
    R = fma(R0, fma(b, R0, np.float32(1.0)), R0)
    Q0 = fma(a, R, np.float32(0.0))
    Q = fma(R, fma(b, Q0, a), Q0)
    assert np.array_equal(Q.view(np.int32), (-q).view(np.int32))
