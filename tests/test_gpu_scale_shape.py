"""Parity at the shapes bench.py measures ("prove parity where you benchmark"): d = 128, 1 M items,
> 32768 eval users (two user blocks: the second one reuses the prepared item operands), ordered pass A at the default stride, 9+ item
splits; and the fused lazy-Adam step kernel at B = 2^20 distinct users against the dense sweep.
Reference: MF/train_new_api.py:594-612 (scoring), MF/model_api.py:102-121,83 (step + Adam)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pda():
    import pda_b200
    assert pda_b200.load().pda_device_count() >= 1
    return pda_b200


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("tables", ["fitted", "near_init"])
def test_eval_at_bench_shape_matches_oracle_and_exact_backend(pda, c_oracle, tables):
    """40 000 users x 1 000 000 items, d = 128, condition: the tcgen05 filter path == the exact CUDA-core backend on
    ALL rows (ids and score bits), and == the C oracle on 384 rows spread over both user blocks."""
    from helpers import synth_interactions
    from oracle import pda_oracle as po
    n_users, n_items, d, K = 40000, 1_000_000, 128, 50
    rng = np.random.default_rng(77)
    if tables == "fitted":
        U = (rng.standard_normal((n_users, d), dtype=np.float32) / np.float32(np.sqrt(d)))
        I = (rng.standard_normal((n_items, d), dtype=np.float32) / np.float32(np.sqrt(d)))
        I *= ((0.2 + rng.random(n_items) ** 4) * 3.0).astype(np.float32)[:, None]
        pop = (rng.random(n_items) ** 6).astype(np.float32)
    else:   # a model a few Adam steps from Xavier init: |s| << 1, the ranking is almost the popularity ranking
        U = (rng.uniform(-1, 1, (n_users, d)) * 7e-4).astype(np.float32)
        I = (rng.uniform(-1, 1, (n_items, d)) * 2.4e-3).astype(np.float32)
        pop = (rng.random(n_items) ** 0.16).astype(np.float32)
    uid, iid, t = synth_interactions(n_users, n_items, 24, 9, seed=3)
    indptr, items, _ = po.build_csr(n_users, uid, iid, t)
    m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=64, init=False)
    m.set_table("user_embedding", U); m.set_table("item_embedding", I)
    m.set_train_csr(indptr, items)
    users = np.arange(n_users, dtype=np.int32)
    ids, sc = m.do_recommendation(users, None, "condition", pos_pop=pop, K=K, backend="tensor", return_scores=True)
    st = m.tc_last_stats()
    plan = np.zeros(24, dtype=np.int64)
    assert pda.load().pda_tc_plan_host(32768, n_items, d, K, plan.ctypes.data) == 0
    assert plan[5] == 1 and plan[7] == 12 and plan[11] >= 9, plan      # ordered pass A, stride 12, >= 9 item splits
    eid, esc = m.do_recommendation(users, None, "condition", pos_pop=pop, K=K, backend="exact", return_scores=True)
    assert np.array_equal(ids, eid), (tables, st, int((ids != eid).any(axis=1).sum()))
    assert np.array_equal(bits(sc), bits(esc))
    sel = np.unique(np.concatenate([np.linspace(0, 32767, 192), np.linspace(32768, n_users - 1, 192)]).astype(np.int32))
    rid, rsc = c_oracle.recommend(U, I, sel, "condition", K, indptr, items, pop=pop)
    assert np.array_equal(ids[sel], rid), (tables, st)
    assert np.array_equal(bits(sc[sel]), bits(rsc))
    print(tables, "filter stats:", st)
    assert st["rows_exact_fallback"] <= 0.05 * st["rows"], st
    m.close()


@pytest.mark.timeout(900)
def test_fused_step_at_bench_batch_matches_dense_sweep(pda, monkeypatch):
    """B = 2^20 distinct users out of 2^21 + 77, 300 k items, d = 128, device sampler: the pipelined fused kernel (lazy
    user table, exact replay) vs the dense sweep vs the register-gather fused kernel.  After ONE step the user table and
    its Adam slots are bit-identical in all three (the item table differs only by the order of the fp32 atomics); after 5
    steps (rows replay 1..4 skipped steps) everything agrees to 1e-4 of scale and the losses to 1e-5."""
    import torch
    from pda_b200 import synth
    n_users, n_items, d, B = (1 << 21) + 77, 300_000, 128, 1 << 20
    dev = torch.device("cuda", 0)
    ds = synth.make_synthetic(n_users, n_items, seed=5, device=dev, mean_extra_deg=4.0, min_deg=3)
    P = synth.train_pop_matrix_torch(ds["pop"], 0.16).cpu().numpy()
    ms = {}
    for name, adam in (("pipe", "lazy_users"), ("reg", "lazy_users"), ("dense", "dense")):
        m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=B, lr=1e-2, regs=1e-3, seed=2021, max_batch=B)
        m.set_train_csr_device(ds["indptr"].data_ptr(), ds["items"].data_ptr(), ds["times"].data_ptr(), ds["nnz"],
                               ds["active"].data_ptr(), ds["active"].numel(), unique_times=np.arange(ds["n_stages"] - 1))
        m.set_train_pop(P)
        m.set_adam_mode(adam)
        ms[name] = m

    def step(n0, n):
        for name, m in ms.items():
            monkeypatch.setenv("PDA_STEP_PIPE", "0" if name == "reg" else "1")
            m.train_sampled(2020, 0, n0, n, B)
            m.synchronize()

    step(0, 1)
    ref = {k: ms["dense"].get_table(k) for k in ("user_embedding", "user_m", "user_v")}
    for name in ("pipe", "reg"):
        for k, r in ref.items():
            assert np.array_equal(bits(ms[name].get_table(k)), bits(r)), (name, k)
    l0 = {k: m.read_loss() for k, m in ms.items()}
    assert np.allclose(l0["pipe"], l0["dense"], rtol=1e-5) and np.allclose(l0["reg"], l0["dense"], rtol=1e-5), l0
    del ref
    step(1, 4)
    l = {k: m.read_loss() for k, m in ms.items()}
    assert np.allclose(l["pipe"], l["dense"], rtol=1e-5) and np.allclose(l["reg"], l["dense"], rtol=1e-5), l
    assert ms["pipe"].adam_stats()[1] > 0
    for k in ("user_embedding", "item_embedding", "user_v"):
        r = ms["dense"].get_table(k)
        scale = np.abs(r).max()
        for name in ("pipe", "reg"):
            g = ms[name].get_table(k)
            assert np.abs(g - r).max() <= 1e-4 * scale, (name, k)
            del g
    for m in ms.values():
        m.close()
