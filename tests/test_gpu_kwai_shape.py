"""BASELINE configs[3] at its shape: PD on a Kwai-shaped synthetic set (Kwai itself is not shipped with the reference:
37 663 users x 128 879 items, ~7.0 M train pairs over 9 stages, gamma = 0.16, d = 64, B = 2048).
Training trajectory (device sampler, fused step, Adam) and PD / PDA evaluation (tcgen05 filter) against the CPU oracle
with shared init and batches."""
import numpy as np
import pytest

from helpers import pop_table, synth_interactions

pytestmark = pytest.mark.gpu

N_USERS, N_ITEMS, T = 37663, 128879, 9


@pytest.fixture(scope="module")
def kwai():
    from oracle import pda_oracle as po
    uid, iid, t = synth_interactions(N_USERS, N_ITEMS, 250, T, seed=11, empty_frac=0.002)
    indptr, items, times = po.build_csr(N_USERS, uid, iid, t)
    pop = pop_table(N_ITEMS, T, seed=12)            # [n_items, T + 1], exact zeros like the real pop tables
    return indptr, items, times, pop


def test_kwai_shape_training_and_eval_match_oracle(kwai, c_oracle):
    import pda_b200
    from oracle import pda_oracle as po
    indptr, items, times, pop = kwai
    assert 6.5e6 < len(items) < 7.5e6
    gamma, B, d, n_steps = 0.16, 2048, 64, 40
    P = po.train_pop_matrix(pop, gamma)
    last, lin = po.eval_pops(pop, gamma)
    m = pda_b200.PDAModel(N_USERS, N_ITEMS, d, train="s_condition", batch_size=B, lr=1e-2, regs=1e-3, seed=2021)
    m.set_train_csr(indptr, items, times, unique_times=np.arange(T))
    m.set_train_pop(P)
    ref = c_oracle.CModel(m.get_table("user_embedding"), m.get_table("item_embedding"), 1e-2, 1e-3, B, "s_condition")
    active = np.nonzero(np.diff(indptr) > 0)[0]
    for s in range(n_steps):
        m.train_sampled(2020, 0, s, 1, B)
        got = m.read_loss()
        b = c_oracle.sample_batch(2020, 0, s, B, active, indptr, items, times, N_ITEMS, np.arange(T), P)
        if s < 3:       # sampled indices are bit-exact
            g = m.sample_batch(2020, 0, s, B)
            for k in ("users", "pos", "neg"):
                assert np.array_equal(g[k], b[k]), (s, k)
        want = ref.train_step(b["users"], b["pos"], b["neg"], b["pos_pop"], b["neg_pop"])
        assert np.allclose(got, want, rtol=1e-5, atol=0), (s, got, want)
    U, I = m.get_table("user_embedding"), m.get_table("item_embedding")
    assert np.abs(U - ref.U).max() <= 1e-4 * np.abs(ref.U).max()
    assert np.abs(I - ref.I).max() <= 1e-4 * np.abs(ref.I).max()
    # evaluation on 3000 users against all 128 879 items: same tables -> the oracle scorer must give the same ids
    rng = np.random.default_rng(3)
    users = np.sort(rng.choice(active, 3000, replace=False)).astype(np.int32)
    for rec_type, p in (("main_branch", None), ("condition", last), ("condition", lin)):
        ids, sc = m.do_recommendation(users, None, rec_type, pos_pop=p, K=50, backend="tensor", return_scores=True)
        st = m.tc_last_stats()
        rid, rsc = c_oracle.recommend(U, I, users, rec_type, 50, indptr, items, pop=p)
        assert np.array_equal(ids, rid), (rec_type, st)
        assert np.array_equal(sc.view(np.int32), rsc.view(np.int32)), rec_type
        assert st["rows_exact_fallback"] <= 0.05 * len(users), st
    m.close()
