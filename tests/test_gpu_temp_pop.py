"""GPU parity of BPR(t)-pop (--train temp_pop; MF/model_api.py:300-401) against the numpy oracle."""
import numpy as np
import pytest

from helpers import synth_interactions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pda():
    import pda_b200
    assert pda_b200.load().pda_device_count() >= 1, "no CUDA device visible: GPU tests cannot run"
    return pda_b200


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


def _sync_tables(m, om):
    for name, arr in (("user_embedding", om.U), ("item_embedding", om.I), ("user_temp_bias", om.ub),
                      ("item_temp_bias", om.ib)):
        assert np.array_equal(bits(m.get_table(name)), bits(arr)), name     # same Philox init (table ids 0..3)


@pytest.mark.parametrize("d", [16, 64, 128])
def test_temp_pop_gradients(pda, d):
    from oracle import pda_oracle as po
    rng = np.random.default_rng(d)
    n_users, n_items, T, B = 500, 200, 9, 256
    om = po.OracleTempPopModel(n_users, n_items, d, T, 1e-2, 1e-3, 2048, seed=2021)
    # larger values so every term matters
    om.U = (rng.normal(0, 0.5, om.U.shape)).astype(np.float32); om.I = rng.normal(0, 0.5, om.I.shape).astype(np.float32)
    om.ub = rng.normal(0, 0.5, om.ub.shape).astype(np.float32); om.ib = rng.normal(0, 0.5, om.ib.shape).astype(np.float32)
    m = pda.PDAModel(n_users, n_items, d, train="temp_pop", batch_size=2048, lr=1e-2, regs=1e-3, max_batch=B, temp_num=T)
    m.set_table("user_embedding", om.U); m.set_table("item_embedding", om.I)
    m.set_table("user_temp_bias", om.ub); m.set_table("item_temp_bias", om.ib)
    users = rng.permutation(n_users)[:B].astype(np.int32)
    pos = (rng.random(B) ** 2 * n_items).astype(np.int32)
    neg = rng.integers(0, n_items, B).astype(np.int32)
    temp = rng.integers(0, T, B).astype(np.int32)
    temp[:40] = 0                                               # t == 0: the only stage where the user bias is read
    gU, gI, gub, gib, loss3 = m.gradients_temp(users, pos, neg, temp)
    r = po.temp_pop_forward_backward(om.U, om.I, om.ub, om.ib, users, pos, neg, temp, 1e-3, 2048)
    refU = np.zeros_like(om.U); refU[users] = r["gU_rows"]
    assert np.array_equal(bits(gU), bits(refU))                 # users are distinct: one term per row -> bit-exact
    refI = np.zeros(om.I.shape, np.float64)
    np.add.at(refI, pos, r["gP_rows"].astype(np.float64)); np.add.at(refI, neg, r["gN_rows"].astype(np.float64))
    assert np.abs(gI - refI).max() <= 1e-5 * np.abs(refI).max()
    ref_ub = np.zeros(n_users, np.float32); ref_ub[users] = r["g_ub"]
    # value equality: a saturated sigmoid gives g = 0 -> the oracle row holds -0.0 where the accumulator holds 0 + -0.0 = +0.0
    assert np.array_equal(gub, ref_ub) and np.array_equal(bits(gub[gub != 0]), bits(ref_ub[ref_ub != 0]))
    assert (gub[users[temp != 0]] == 0).all() and np.abs(gub[users[:40]]).max() > 0
    ref_ib = np.zeros(om.ib.shape, np.float64)
    np.add.at(ref_ib, (pos, np.full(B, T)), r["g_pib"].astype(np.float64)); np.add.at(ref_ib, (pos, temp), r["g_pib"].astype(np.float64))
    np.add.at(ref_ib, (neg, np.full(B, T)), r["g_nib"].astype(np.float64)); np.add.at(ref_ib, (neg, temp), r["g_nib"].astype(np.float64))
    assert np.abs(gib - ref_ib).max() <= 1e-5 * np.abs(ref_ib).max()
    assert np.allclose(loss3, [r["loss"], r["mf_loss"], r["reg_loss"]], rtol=1e-5, atol=0)
    m.close()


def test_temp_pop_training_bit_exact_without_duplicates(pda):
    from oracle import pda_oracle as po
    rng = np.random.default_rng(3)
    n_users, n_items, T, B, d = 600, 900, 9, 128, 64
    om = po.OracleTempPopModel(n_users, n_items, d, T, 1e-2, 1e-3, B, seed=2021)
    m = pda.PDAModel(n_users, n_items, d, train="temp_pop", batch_size=B, lr=1e-2, regs=1e-3, temp_num=T, seed=2021)
    _sync_tables(m, om)
    for step in range(5):
        users = rng.permutation(n_users)[:B].astype(np.int32)
        it = rng.permutation(n_items)[:2 * B].astype(np.int32)
        temp = rng.integers(0, T, B).astype(np.int32)
        got = m.train_step(users, it[:B], it[B:], temp.astype(np.float32), np.arange(B, dtype=np.float32))
        want = om.train_step(users, it[:B], it[B:], temp)
        assert np.allclose(got, want, rtol=1e-5, atol=0), (step, got, want)
    _sync_tables(m, om)                                          # all four variables, bit for bit, after 5 Adam steps
    first = 17
    assert np.array_equal(bits(m.temp_item_bias_for_eval(first)), bits(om.item_bias_for_eval(first)))
    m.close()


def test_temp_pop_sampled_training_and_eval(pda, c_oracle):
    """device sampler (stage of the sampled interaction -> b_time) -> fused step -> Adam over 4 variables, then the
    reference's per-batch eval rule (bias factor of the batch's first user)."""
    from oracle import pda_oracle as po
    n_users, n_items, T, B, d = 1500, 400, 9, 256, 32
    uid, iid, t = synth_interactions(n_users, n_items, 10, T, seed=12)
    indptr, items, times = po.build_csr(n_users, uid, iid, t)
    om = po.OracleTempPopModel(n_users, n_items, d, T, 1e-2, 1e-3, B, seed=2021)
    m = pda.PDAModel(n_users, n_items, d, train="temp_pop", batch_size=B, lr=1e-2, regs=1e-3, temp_num=T, seed=2021)
    m.set_train_csr(indptr, items, times, unique_times=np.arange(T))
    active = np.nonzero(np.diff(indptr) > 0)[0]
    m.train_sampled(2020, 0, 0, 8, B)
    got = m.read_loss()
    for s in range(8):
        b = c_oracle.sample_batch(2020, 0, s, B, active, indptr, items, times, n_items, np.arange(T))
        want = om.train_step(b["users"], b["pos"], b["neg"], b["time"])
    assert np.allclose(got, want, rtol=1e-5, atol=0)
    for name, arr in (("user_embedding", om.U), ("item_embedding", om.I), ("user_temp_bias", om.ub), ("item_temp_bias", om.ib)):
        assert np.abs(m.get_table(name) - arr).max() <= 1e-4 * np.abs(arr).max(), name
    users = np.arange(100, 300, dtype=np.int32)
    bias = m.temp_item_bias_for_eval(users[0])
    ids = m.do_recommendation(users, None, "main_branch", K=50, col_bias=bias, backend="exact")
    rid, _ = c_oracle.recommend(m.get_table("user_embedding"), m.get_table("item_embedding"), users, "main_branch", 50,
                                indptr, items, col_bias=bias)
    assert np.array_equal(ids, rid)
    sd = m.state_dict()
    assert "parameter/user_temp_bias" in sd and sd["parameter/item_temp_bias"].shape == (n_items, T + 1)
    m.close()


def test_temp_pop_needs_temp_num_and_stages(pda):
    with pytest.raises(pda.PdaError):
        pda.PDAModel(10, 10, 8, train="temp_pop")               # temp_num missing
    m = pda.PDAModel(50, 40, 8, train="temp_pop", batch_size=16, temp_num=3)
    m.set_train_csr(np.arange(51, dtype=np.int64), np.arange(50, dtype=np.int32) % 40)    # no stage labels
    with pytest.raises(pda.PdaError):
        m.sample_batch(1, 0, 0)
    with pytest.raises(pda.PdaError):
        m.train_step(np.arange(16), np.arange(16), np.arange(16))                           # stage array missing
    m.close()
