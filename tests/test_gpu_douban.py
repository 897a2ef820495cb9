"""End-to-end on the shipped Douban set (BASELINE configs[1], configs[2]): the MF/train_new_api.py CLI, and
trajectory parity of PD training + PDA evaluation against the CPU oracle with shared init and batches.
Needs the parsed caches of tools/stage_douban.py under ./data/douban (they travel with gpurun)."""
import os
import re
import subprocess
import sys
from types import SimpleNamespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOUBAN = os.path.join(ROOT, "data", "douban")
needs_douban = pytest.mark.skipif(not os.path.exists(os.path.join(DOUBAN, "pda_cache_Data2.npz")),
                                  reason="data/douban caches absent (run tools/stage_douban.py where the reference zip exists)")


@pytest.fixture(scope="module")
def douban():
    from pda_b200 import data as D, popularity as P
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        args = SimpleNamespace(dataset="douban", batch_size=2048, model="mf", data_path="./data/")
        d = D.Data2(args)
        pop = P.load_popularity(args)
    finally:
        os.chdir(cwd)
    return d, pop


@needs_douban
def test_douban_shapes(douban):
    d, pop = douban
    assert (d.n_users, d.n_items, d.n_train, d.n_valid, d.n_test) == (47890, 26047, 6625965, 168716, 379537)
    assert len(d.valid_user_list) == 6847 and len(d.test_user_list) == 15974
    assert sorted(d.unique_times) == list(range(9)) and pop.shape == (26047, 10)
    assert d.n_train // 2048 + 1 == 3236


@needs_douban
def test_douban_pd_training_and_pda_eval_match_oracle(douban, c_oracle):
    """PD (gamma = 0.22, d = 64, B = 2048, lr 1e-2, regs 1e-3), 200 steps from the shared Philox init with the shared
    Philox batches, then PD / PDA evaluation on the valid users: per-step losses within 1e-5, Recall@20 / NDCG@20
    within 1e-4 (north_star), top-50 ids compared row by row."""
    import pda_b200
    from oracle import pda_oracle as po
    d, pop = douban
    gamma, B, dim, n_steps = 0.22, 2048, 64, 200
    P = po.train_pop_matrix(pop, gamma)
    last, lin = po.eval_pops(pop, gamma)
    m = pda_b200.PDAModel(d.n_users, d.n_items, dim, train="s_condition", batch_size=B, lr=1e-2, regs=1e-3, seed=2021)
    m.set_train_csr(d.train_indptr, d.train_items, d.train_times, unique_times=d.unique_times)
    m.set_train_pop(P)
    ref = c_oracle.CModel(m.get_table("user_embedding"), m.get_table("item_embedding"), 1e-2, 1e-3, B, "s_condition")
    active = np.nonzero(np.diff(d.train_indptr) > 0)[0]
    for s in range(n_steps):
        m.train_sampled(2020, 0, s, 1, B)
        if s < 50 or s == n_steps - 1:
            got = m.read_loss()
        b = c_oracle.sample_batch(2020, 0, s, B, active, d.train_indptr, d.train_items, d.train_times, d.n_items,
                                  d.unique_times, P)
        want = ref.train_step(b["users"], b["pos"], b["neg"], b["pos_pop"], b["neg_pop"])
        if s < 50 or s == n_steps - 1:
            assert np.allclose(got, want, rtol=1e-5, atol=0), (s, got, want)
    U, I = m.get_table("user_embedding"), m.get_table("item_embedding")
    assert np.abs(U - ref.U).max() <= 1e-4 * np.abs(ref.U).max()
    assert np.abs(I - ref.I).max() <= 1e-4 * np.abs(ref.I).max()
    users = np.asarray(d.valid_user_list.keys(), dtype=np.int32)
    for rec_type, p in (("main_branch", None), ("condition", last), ("condition", lin)):
        ids = m.do_recommendation(users, None, rec_type, pos_pop=p, K=50)
        st = m.tc_last_stats()
        # the filter carries the load on a model that has started to fit (heavy users' train items loosen tau: the candidate
        # capacity must hold them) -- at most a handful of rows may need the exact kernel
        assert st["rows_exact_fallback"] <= 0.01 * len(users), (rec_type, st)
        got = m.metrics_sum(ids, users, d.valid_indptr, d.valid_items, [20, 50])
        rid, _ = c_oracle.recommend(ref.U, ref.I, users, rec_type, 50, d.train_indptr, d.train_items, pop=p)
        want = c_oracle.metrics_sum(rid, users, d.valid_indptr, d.valid_items, [20, 50])
        for k in ("recall", "ndcg", "precision", "hit_ratio"):
            assert np.abs(got[k] - want[k]).max() / len(users) <= 1e-4, (rec_type, k, got[k] / len(users), want[k] / len(users))
        # same tables -> same ids: score the GPU's tables with the oracle scorer
        rid2, _ = c_oracle.recommend(U, I, users, rec_type, 50, d.train_indptr, d.train_items, pop=p)
        assert np.array_equal(ids, rid2), rec_type
    m.close()


@needs_douban
def test_cli_runs_pd_on_douban(tmp_path):
    """python MF/train_new_api.py ... exactly as README.md:69 of the reference spells it (2 epochs)."""
    cmd = [sys.executable, "-u", "MF/train_new_api.py", "--dataset", "douban", "--epoch", "2", "--save_flag", "0",
           "--log_interval", "1", "--start", "0", "--end", "10", "--step", "1", "--batch_size", "2048", "--lr", "1e-2",
           "--train", "s_condition", "--test", "s_condition", "--saveID", "s_condition", "--cuda", "0", "--regs", "1e-3",
           "--valid_set", "valid", "--pop_exp", "0.22", "--save_dir", str(tmp_path) + "/", "--Ks", "[20,50]"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert "batch_num: 3236" in out and "-------    running PD & PDA model  ----------------" in out
    ep = re.findall(r"Epoch (\d+) \[[\d.]+s\]: train==\[([\d.]+)=([\d.]+) \+ ([\d.]+)\]", out)
    assert [e[0] for e in ep] == ["0", "1"]
    assert float(ep[1][1]) < float(ep[0][1]) < 0.6932            # the loss falls from ~log 2
    rec = re.findall(r"recall=\[([\d.]+), ([\d.]+)\]", out)
    assert len(rec) >= 10 and all(0.0 <= float(a) <= float(b) <= 1.0 for a, b in rec)
    # per evaluation the driver prints PD (main_branch), PDA with last-stage pop, PDA with linear pop
    assert float(rec[4][0]) > 0.02                                # PDA after 2 epochs: far above chance (50/26047)
    assert float(rec[4][0]) > float(rec[3][0])                    # injecting popularity helps on Douban (the paper's point)
    assert "training and testing end!!!!" in out
    ck = [f for _, _, fs in os.walk(tmp_path) for f in fs]
    assert "best_ckpt.ckpt.npz" in ck and "best_main_ckpt.ckpt.npz" in ck


@needs_douban
@pytest.mark.parametrize("train", ["normal", "condition", "temp_pop"])
def test_cli_other_models_run_on_douban(tmp_path, train):
    """BPRMF (+ the BPRMF-A gamma~ line search), PDG and BPR(t)-pop through the same CLI, 1 epoch each."""
    cmd = [sys.executable, "-u", "MF/train_new_api.py", "--dataset", "douban", "--epoch", "1", "--save_flag", "0",
           "--log_interval", "1", "--batch_size", "2048", "--lr", "1e-2", "--train", train, "--test", train, "--saveID", "t",
           "--cuda", "0", "--regs", "1e-3", "--valid_set", "valid", "--pop_exp", "0.22", "--save_dir", str(tmp_path) + "/",
           "--Ks", "[20,50]"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert "training and testing end!!!!" in out and "batch_num: 3236" in out
    rec = [float(a) for a, _ in re.findall(r"recall=\[([\d.]+), ([\d.]+)\]", out)]
    assert len(rec) >= 3 and max(rec) > 0.01
    if train == "normal":
        assert "best expo:" in out and "BPRMF-A with injecting last stage pop(best gamma)" in out
        expo = [float(x) for x in re.findall(r"expo: ([\d.]+) best expo", out)]
        assert expo[0] == 0.04 and len(expo) >= 5            # 0.04, 0.06, ... until 5 non-improvements
    if train == "temp_pop":
        assert "running temproal pop MF" in out
