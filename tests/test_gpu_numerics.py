"""Hardware evidence for the bit-exactness claims the exact Adam replay and the tcgen05 filter rest on.

1. The straight-line MUFU + FFMA refinements of pda_common.cuh (no per-element slow-path branch) must return the bits of
   __fsqrt_rn / __fdiv_rn for EVERY operand inside their guarded ranges: an exhaustive sweep of the sqrt range (1.17 G
   fp32 values), >= 2^32 random in-range quotients, and the packed zero-gradient / gradient Adam steps as they are used
   (f32x2 instructions, the sign-mirrored quotient) against the generic separately rounded form.
2. The certificate of the tensor-core filter assumes the MMA's fp32 accumulation loses at most cB x magnitude
   (DESIGN.md 5.4).  Adversarial accumulators: operands exactly representable in bf16 (so the operand-rounding term of the
   bound is zero), all of one sign (no cancellation), exponents spread over 2^-12 .. 2^12, d = 64 / 128 -- the raw TMEM
   accumulators must stay within cB x (|u| |w_j| + |x_j|) of the fp64 sum.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pda():
    import pda_b200
    assert pda_b200.load().pda_device_count() >= 1
    return pda_b200


def _numerics(pda, kind, a, b, per_thread):
    out = np.zeros(5, dtype=np.uint64)
    rc = pda.load().pda_debug_numerics(kind, a, b, per_thread, C.c_void_p(out.ctypes.data))
    assert rc == 0
    return int(out[0]), int(out[1]), [hex(int(x)) for x in out[2:]]


def test_sqrt_refinement_exhaustive(pda):
    """every fp32 in [2^-100, 2^40]: sqrt_rn_inrange(x) == __fsqrt_rn(x)"""
    lo, hi = (127 - 100) << 23, (127 + 40) << 23
    n, bad, first = _numerics(pda, 3, lo, hi, 0)
    assert n == hi - lo + 1
    assert bad == 0, (bad, first)


def test_div_refinement_random_2_to_33(pda):
    """2^33 random pairs, |a| in [2^-100, 2^60), b in [2^-27, 2^21): div_rn_inrange(a, b) == __fdiv_rn(a, b)"""
    threads = 148 * 8 * 256
    per_thread = (1 << 33) // (2 * threads) + 1
    n, bad, first = _numerics(pda, 0, 12345, 0, per_thread)
    assert n >= 1 << 33
    assert bad == 0, (bad, first)


@pytest.mark.parametrize("kind", [1, 2, 4])
def test_packed_adam_steps_equal_generic_form(pda, kind):
    """2^31 random in-range elements through zero_grad_step4_unguarded (kind 1) / lazy_grad_step4 (kind 2) / three
    steps of the negated-v form zero_grad_step4_nv (kind 4) vs lazy_zero_grad_step / lazy_grad_step: w, m, v bit-identical"""
    threads = 148 * 8 * 256
    per_thread = (1 << 31) // (4 * threads) + 1
    n, bad, first = _numerics(pda, kind, 777 + kind, 0, per_thread)
    assert n >= 1 << 31
    assert bad == 0, (bad, first)


def _bf16_exact(rng, shape, e_lo, e_hi):
    """positive floats with 8 significant bits (exactly representable in bf16) and exponents uniform in [e_lo, e_hi]"""
    mant = rng.integers(128, 256, shape).astype(np.float64) / 128.0
    return (mant * 2.0 ** rng.integers(e_lo, e_hi + 1, shape)).astype(np.float32)


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("kind", ["main_branch", "condition"])
@pytest.mark.parametrize("spread", [0, 6, 12])
def test_accumulation_term_of_the_bound_adversarial(pda, d, kind, spread):
    """same-sign, bf16-exact operands with exponents spread over 2^-spread .. 2^spread: the only error left is the tensor
    core's accumulation (+ the 3 x bf16 split of x_j): it must stay below cB (|u| |w_j| + |x_j|)."""
    rng = np.random.default_rng(100 + d + spread)
    n_users, n_items = 256, 8192
    U = _bf16_exact(rng, (n_users, d), -spread, spread)
    I = _bf16_exact(rng, (n_items, d), -spread, spread)
    if spread == 12:      # one dominant product per row/column pair + many tiny ones: alignment truncation in the adder
        U[:, 0] *= 2.0 ** 6; I[:, 0] *= 2.0 ** 6
    pop = (2.0 ** rng.integers(-6, 1, n_items)).astype(np.float32)       # powers of two: pop_j * i_j stays bf16-exact
    m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=64, init=False)
    m.set_table("user_embedding", U); m.set_table("item_embedding", I)
    users = np.arange(n_users, dtype=np.int32)
    S = U.astype(np.float64) @ I.astype(np.float64).T
    un = np.linalg.norm(U.astype(np.float64), axis=1)[:, None]
    if kind == "condition":
        v, (cAB, cB) = m.tc_debug_dense(users, "condition", pos_pop=pop)
        want = (S + 1.0) * pop[None, :].astype(np.float64)
        wn = np.linalg.norm(I.astype(np.float64) * pop[:, None], axis=1)[None, :]
        xa = np.abs(pop)[None, :].astype(np.float64)
    else:
        v, (cAB, cB) = m.tc_debug_dense(users, "main_branch")
        want, wn, xa = S, np.linalg.norm(I.astype(np.float64), axis=1)[None, :], np.zeros((1, n_items))
    err = np.abs(v.astype(np.float64) - want)
    E_acc = cB * (un * wn + xa)
    ratio = float((err / E_acc).max())
    print("d=%d %s spread=%d: max |err| / (cB (|u||w| + |x|)) = %.4f; max rel err vs |want| = %.3e" %
          (d, kind, spread, ratio, float((err / np.abs(want)).max())))
    assert ratio <= 1.0, ratio
    m.close()
