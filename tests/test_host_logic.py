"""Host-side logic without a GPU: flag parsing, the data loaders on files in the reference's formats, popularity
preparation, the evaluation driver and early stopping (with a stand-in recommender where a GPU would be needed)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------
# flags (MF/parse.py)
# ---------------------------------------------------------------------------------------------
def _parser():
    sys.path.insert(0, os.path.join(ROOT, "MF"))
    try:
        import importlib
        return importlib.import_module("parse")
    finally:
        sys.path.pop(0)


def test_reference_command_lines_parse():
    p = _parser()
    # README.md:69 of the reference (PD / PDA) and :41 (BPRMF), flags the reference marks "not used" included
    a = p.parse_args("--dataset douban --epoch 2000 --save_flag 0 --log_interval 5 --start 0 --end 10 --step 1 "
                     "--batch_size 2048 --lr 1e-2 --train s_condition --test s_condition --saveID s_condition --cuda 0 "
                     "--regs 1e-3 --valid_set valid --pop_exp 0.22 --save_dir /tmp/x/ --Ks [20,50]".split())
    assert (a.dataset, a.train, a.test, a.valid_set) == ("douban", "s_condition", "s_condition", "valid")
    assert (a.batch_size, a.lr, a.regs, a.pop_exp, a.log_interval) == (2048, 1e-2, 1e-3, 0.22, 5)
    assert eval(a.Ks) == [20, 50] and a.cuda == "0" and a.step == 1
    d = p.parse_args([])
    assert (d.dataset, d.train, d.embed_size, d.batch_size, d.lr, d.regs, d.Ks, d.epoch) == \
        ("kwai", "normal", 64, 1024, 1e-3, 1e-5, "[20]", 400)                       # parse.py defaults
    assert d.early_stop == 1 and d.log_interval == 10 and d.data_path == "./data/" and d.model == "mf"
    with pytest.raises(SystemExit):
        p.parse_args(["--no_such_flag", "1"])


# ---------------------------------------------------------------------------------------------
# data loaders (MF/load_data.py formats)
# ---------------------------------------------------------------------------------------------
def _write_dataset(root, name="toy", with_time=True):
    rng = np.random.default_rng(0)
    n_users, n_items, T = 40, 30, 4
    d = os.path.join(root, "data", name)
    os.makedirs(d)
    train = {}
    rows = []
    for u in range(n_users):
        if u == 7:
            continue                                     # a user without training data
        its = rng.permutation(n_items)[: rng.integers(1, 9)]
        train[u] = [int(i) for i in its]
        for i in its:
            rows.append((u, int(i), int(rng.integers(0, T)), 5))
    rng.shuffle(rows)
    if with_time:
        with open(os.path.join(d, "train_with_time.txt"), "w") as f:
            for r in rows:
                f.write("%d %d %d %d\n" % r)
    with open(os.path.join(d, "train.txt"), "w") as f:
        for u, its in train.items():
            f.write(" ".join(str(x) for x in [u] + its) + "\n")
    valid = {u: [int(x) for x in rng.permutation(n_items)[:3]] for u in (3, 1, 20)}
    test = {u: [int(x) for x in rng.permutation(n_items)[:4]] for u in (39, 2, 11, 5)}
    for fn, dd in (("valid.txt", valid), ("test.txt", test)):
        with open(os.path.join(d, fn), "w") as f:
            for u, its in dd.items():
                f.write(" ".join(str(x) for x in [u] + its) + "\n")
            f.write("9\n")                                # a user line without items is skipped (load_data.py:60)
    pop = rng.random((n_items, T + 1))
    pop[rng.random(pop.shape) < 0.2] = 0.0
    with open(os.path.join(d, "item_pop_seq_ori2.txt"), "w") as f:
        for i in range(n_items):
            f.write(" ".join([str(i)] + [repr(float(x)) for x in pop[i]]) + "\n")
    return d, train, rows, valid, test, pop


def test_data2_loader_builds_sorted_csr(tmp_path, monkeypatch):
    from pda_b200 import data as D
    d, train, rows, valid, test, pop = _write_dataset(str(tmp_path))
    monkeypatch.chdir(tmp_path)
    a = SimpleNamespace(dataset="toy", batch_size=8, model="mf", data_path="./data/")
    ds = D.Data2(a)
    assert (ds.n_users, ds.n_items) == (40, 30) and ds.n_train == len(rows)
    assert ds.n_valid == 9 and ds.n_test == 16
    for u in range(40):
        seg = ds.train_items[ds.train_indptr[u]:ds.train_indptr[u + 1]]
        assert sorted(train.get(u, [])) == seg.tolist()
        assert (np.diff(seg) >= 0).all()
    lut = {(u, i): t for u, i, t, _ in rows}
    for u in (0, 5, 39):
        for i, t in zip(ds.train_user_list[u], ds.train_user_list_time[u]):
            assert lut[(u, i)] == t                       # the stage travels with its interaction
    first_seen = []
    for r in rows:
        if r[2] not in first_seen:
            first_seen.append(r[2])
    assert ds.unique_times == first_seen                   # Series.unique(): order of appearance (load_data.py:629)
    assert 7 not in ds.train_user_list and ds.train_user_list[7] == []
    assert ds.valid_user_list.keys() == [3, 1, 20] and ds.test_user_list.keys() == [39, 2, 11, 5]   # file order
    assert ds.test_user_list[11] == test[11] and ds.valid_user_list[20] == valid[20]
    # second construction comes from the cache and is identical
    assert os.path.exists(os.path.join(d, "pda_cache_Data2.npz"))
    ds2 = D.Data2(a)
    assert np.array_equal(ds.train_items, ds2.train_items) and np.array_equal(ds.train_times, ds2.train_times)
    assert ds2.unique_times == ds.unique_times and ds2.test_user_list.keys() == [39, 2, 11, 5]
    # cache-only directory (what a GPU box receives)
    for f in ("train_with_time.txt", "train.txt", "valid.txt", "test.txt"):
        os.remove(os.path.join(d, f))
    ds3 = D.Data2(a)
    assert np.array_equal(ds3.train_indptr, ds.train_indptr) and ds3.n_test == 16


def test_data_loader_train_txt(tmp_path, monkeypatch):
    from pda_b200 import data as D
    d, train, rows, valid, test, pop = _write_dataset(str(tmp_path), with_time=False)
    monkeypatch.chdir(tmp_path)
    ds = D.Data(SimpleNamespace(dataset="toy", batch_size=8, model="mf", data_path="./data/"))
    assert ds.train_times is None and ds.unique_times == [] and ds.n_train == sum(len(v) for v in train.values())
    assert sorted(ds.train_user_list[3]) == sorted(train[3])
    with pytest.raises(FileNotFoundError):
        D.Data(SimpleNamespace(dataset="nope", batch_size=8, model="mf", data_path="./data/"))
    with pytest.raises(NotImplementedError):
        D.Data(SimpleNamespace(dataset="toy", batch_size=8, model="CausalE", data_path="./data/"))


def test_popularity_preparation(tmp_path, monkeypatch, capsys):
    from oracle import pda_oracle as po
    from pda_b200 import popularity as P
    d, train, rows, valid, test, pop = _write_dataset(str(tmp_path))
    monkeypatch.chdir(tmp_path)
    a = SimpleNamespace(dataset="toy", data_path="./data/")
    got = P.load_popularity(a)
    assert np.array_equal(got, pop)                        # repr() round-trips float64
    assert "popularity used: ./data/toy/item_pop_seq_ori2.txt" in capsys.readouterr().out
    last, lin = P.eval_popularities(got.copy(), 0.22)
    o_last, o_lin = po.eval_pops(pop, 0.22)
    assert np.array_equal(last.astype(np.float32), o_last) and np.array_equal(lin.astype(np.float32), o_lin)
    train_pop = np.power(P.get_popularity_from_load(got), 0.22)
    assert np.array_equal(train_pop.astype(np.float32), po.train_pop_matrix(pop, 0.22))
    assert train_pop.shape == (30, 4)
    # BPRMF-A quirk B.5: the clip masks come from the powered array -> the raw linear prediction is NOT clipped
    lo, li = P.bprmf_a_popularities(got.copy(), lin)
    raw = pop[:, -2] + 0.5 * (pop[:, -2] - pop[:, -3])
    assert np.array_equal(lo, pop[:, -2]) and np.array_equal(li, raw)


def test_stage_file_formula_matches_oracle(tmp_path):
    from oracle import pda_oracle as po
    from pda_b200 import popularity as P
    rng = np.random.default_rng(1)
    n_items, T = 25, 3
    counts = rng.integers(0, 6, (T, n_items))
    for t in range(T):
        with open(os.path.join(tmp_path, "t_%d.txt" % t), "w") as f:
            for i in range(n_items):
                if counts[t, i]:
                    f.write(" ".join([str(i)] + ["1"] * int(counts[t, i])) + "\n")
    got = P.pop_table_from_stage_files(str(tmp_path), T, n_items)
    assert np.abs(got - po.pop_table_from_stage_counts(counts)).max() <= 1e-15


# ---------------------------------------------------------------------------------------------
# evaluation driver + early stop with a stand-in recommender
# ---------------------------------------------------------------------------------------------
class _FakeRecommender:
    """ranks items by a fixed score table on the host; metrics through the oracle (the GPU model's role)."""

    def __init__(self, scores, train_indptr, train_items, c_oracle):
        self.scores, self.ip, self.it, self.co = scores, train_indptr, train_items, c_oracle
        self.calls = []

    def do_recommendation(self, sess, batch_users, items, rec_type, pos_pop=None, sparse_cliked_matrix=None):
        self.calls.append((len(batch_users), rec_type, pos_pop is not None))
        Y = self.scores[np.asarray(batch_users)].copy()
        if pos_pop is not None:
            Y = Y * np.asarray(pos_pop)[None, :]
        for r, u in enumerate(batch_users):
            Y[r, self.it[self.ip[u]:self.ip[u + 1]]] = -np.inf
        from oracle import pda_oracle as po
        return po.topk_ids(Y, 50)

    def metrics_sum(self, ids, eval_users, truth_indptr, truth_items, Ks):
        return self.co.metrics_sum(ids, eval_users, truth_indptr, truth_items, Ks)


def test_evaluation_class_matches_reference_protocol(tmp_path, monkeypatch, c_oracle):
    from oracle import pda_oracle as po
    from pda_b200 import data as D
    from pda_b200.evaluation import evaluation
    _write_dataset(str(tmp_path))
    monkeypatch.chdir(tmp_path)
    # widen the item space so a top-50 exists
    ds = D.Data2(SimpleNamespace(dataset="toy", batch_size=8, model="mf", data_path="./data/"))
    rng = np.random.default_rng(5)
    ds.n_items = 80
    scores = rng.normal(size=(ds.n_users, ds.n_items)).astype(np.float32)
    fake = _FakeRecommender(scores, ds.train_indptr, ds.train_items, c_oracle)
    ev = evaluation(ds, [20, 50], batch_size=3)
    ev.set_evaluate_obj_pre('test')
    assert ev.tot_user == 4 and [len(b) for b in ev.list_batch_user] == [3, 1]
    ret = ev.eval(fake, None, 'main_branch')
    # the reference protocol, literally: per user get_performance on the masked ranking, mean over eval users
    want = {k: np.zeros(2) for k in ret}
    for u in ds.test_user_list.keys():
        y = scores[u].copy()
        y[ds.train_user_list[u]] = -np.inf
        one = po.get_performance(ds.test_user_list[u], po.topk_ids(y[None, :], 50)[0], [20, 50])
        for k in want:
            want[k] += one[k] / 4
    for k in want:
        assert np.allclose(ret[k], want[k], atol=1e-12), k
    assert fake.calls == [(4, 'main_branch', False)]           # one call for all eval users
    fake.needs_reference_eval_batches = True                   # BPR(t)-pop: reference batching is observable
    ev.set_testing_popularity(np.ones(ds.n_items))
    ret2 = ev.eval(fake, None, 'main_with_pop')
    assert [c[0] for c in fake.calls[1:]] == [3, 1] and all(c[2] for c in fake.calls[1:])
    for k in want:
        assert np.allclose(ret2[k], want[k], atol=1e-12)
    ev.set_evaluate_obj_pre('valid')
    assert ev.all_users.tolist() == [3, 1, 20]


def test_early_stop_rule():
    from pda_b200.driver import early_stop
    cfg = dict(best_hr=0, best_ndcg=0, best_recall=0, best_pre=0, best_epoch=0)
    step = 0
    seq = [0.10, 0.12, 0.12, 0.11, 0.11, 0.10]
    stops = []
    for ep, r in enumerate(seq):
        cfg, step, stop = early_stop(r, r, r, r, ep, cfg, step, flag_step=3)
        stops.append(stop)
    assert cfg['best_epoch'] == 2 and cfg['best_recall'] == 0.12       # a tie counts as an improvement (>=)
    assert stops == [False, False, False, False, False, True]


# ---------------------------------------------------------------------------------------------
# launch plan of the tensor-core eval (pure host arithmetic of libpda_b200: runs without a GPU)
# ---------------------------------------------------------------------------------------------
PLAN_KEYS = ("M_pad", "N_pad", "n_tiles", "mr", "ts", "ordered", "n_sel", "se", "cw", "n_c", "n_valid", "splits",
             "tiles_per_split", "n_seg", "seg_cap", "rc", "total", "o_Ib", "o_Ub", "o_cmax", "o_cand", "o_clist", "o_work",
             "o_nwork")


def _plan(M, N, d, K=50):
    import pda_b200
    from pda_b200._lib import check, ptr
    out = np.zeros(24, dtype=np.int64)
    check(pda_b200.load().pda_tc_plan_host(M, N, d, K, ptr(out)))
    return dict(zip(PLAN_KEYS, (int(x) for x in out)))


@pytest.mark.parametrize("N", [4096, 5000, 26047, 128879, 131073, 300000, 1_000_000, 8_000_000])
@pytest.mark.parametrize("M", [1, 300, 6847, 15974, 32768])
@pytest.mark.parametrize("d", [64, 128])
def test_tensor_eval_plan_invariants(M, N, d):
    p = _plan(M, N, d)
    rows = 128 * p["mr"]
    assert p["M_pad"] % rows == 0 and 0 <= p["M_pad"] - M < rows
    assert p["N_pad"] % 128 == 0 and 0 <= p["N_pad"] - N < 128 and p["n_tiles"] == p["N_pad"] // 128
    cpt = 128 // p["cw"]
    assert p["cw"] in (32, 64) and p["se"] >= 1
    assert p["n_sel"] == -(-p["n_tiles"] // p["se"]) and p["n_valid"] == p["n_sel"] * cpt
    assert p["n_c"] % 4 == 0 and p["n_valid"] <= p["n_c"] < p["n_valid"] + 4 and p["n_c"] <= 4096 + 3      # keys fit the selection kernel
    assert (p["se"] == 1) == (p["n_tiles"] * 4 <= 4096) and p["ordered"] == (1 if p["se"] > 1 else 0)
    assert p["splits"] >= 1 and p["splits"] * p["tiles_per_split"] >= p["n_tiles"] > (p["splits"] - 1) * p["tiles_per_split"]
    assert p["n_seg"] == 2 * p["splits"] and p["seg_cap"] >= 64 and p["seg_cap"] % 32 == 0
    assert p["rc"] in (512, 1024, 2048) and p["rc"] // 32 <= 64                                            # work item = row << 6 | round
    # whole waves of CTAs (one CTA per SM) wherever the item set is long enough to split freely
    ctas = (p["M_pad"] // rows) * p["splits"]
    if p["n_tiles"] // p["se"] >= 200 and M >= 6847:
        assert ctas / 148 / -(-ctas // 148) >= 0.9, ctas
    offs = [p[k] for k in ("o_Ib", "o_Ub", "o_cmax", "o_cand", "o_clist", "o_work", "o_nwork")]
    # after the work counter: the partial top-K lists of the few-row exact fallback (1024 rows x 64 item splits x K, values + ids)
    part = 2 * ((1024 * 64 * 50 * 4 + 255) // 256 * 256)
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs) and offs[-1] + 256 + part == p["total"]
    assert p["o_Ub"] - p["o_Ib"] >= p["N_pad"] * d * 2                    # item operands first: they outlive a user block
    assert p["total"] < 16 << 30                                          # per 32768-user block, well inside 180 GB


def test_tensor_eval_plan_item_side_is_independent_of_the_user_block():
    a, b = _plan(32768, 1_000_000, 128), _plan(513, 1_000_000, 128)
    for k in ("N_pad", "n_tiles", "se", "cw", "n_sel", "n_valid", "o_Ib", "o_Ub"):
        assert a[k] == b[k], k
    with pytest.raises(Exception, match="tensor-core eval needs"):
        _plan(100, 5000, 32)
