"""Generates the golden fixtures of tests/golden/ from the REFERENCE's own runnable code.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
  metrics_used_metric.json  <- MF/used_metric.py:get_performance, imported unmodified (np.float shim)
  cpp_evaluator.npz         <- evaluator/backend/cpp/include/{evaluate,metric}.h and
                               util/cython/include/arg_topk.h compiled to oracle/_ref/libref_eval.so
  douban_pop_slice.npz      <- data/douban/douban.zip: t_k.txt counts + the shipped item_pop_seq_ori2.txt
  oracle_step_eval.npz      <- NOT from the reference (its TF1 graph cannot run): the numpy oracle's own output on a
                               256 x 512 slice (3 PD steps without duplicate items + PD / PDA top-50), SURVEY 8c --
                               pins the oracle against drift; the C oracle and the CUDA path must reproduce it bit for bit
The GPU box has no /root/reference; tests only read the committed fixtures.
"""
import io
import json
import os
import sys
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def gen_metrics():
    np.float = float  # removed in numpy >= 1.24; used_metric.py:36,48 still uses it
    sys.path.insert(0, os.path.join(REF, "MF"))
    import used_metric
    rng = np.random.default_rng(20211)
    cases = []
    for c in range(60):
        n_items = int(rng.integers(60, 400))
        n_truth = int(rng.integers(1, 70))
        truth = sorted(int(x) for x in rng.permutation(n_items)[:n_truth])
        ids = [int(x) for x in rng.permutation(n_items)[:50]]
        if c % 5 == 0:   # force hits at the head
            ids[: min(5, n_truth)] = truth[: min(5, n_truth)]
        Ks = [[20, 50], [1, 5, 10], [50], [3, 20]][c % 4]
        r = used_metric.get_performance(truth, ids, Ks)
        cases.append(dict(truth=truth, ids=ids, Ks=Ks, **{k: [float(x) for x in v] for k, v in r.items()}))
    # the README-style toy case quoted in SURVEY 8c
    r = used_metric.get_performance([1, 5], [5, 2, 3, 1], [2, 4])
    cases.append(dict(truth=[1, 5], ids=[5, 2, 3, 1], Ks=[2, 4], **{k: [float(x) for x in v] for k, v in r.items()}))
    json.dump(cases, open(os.path.join(HERE, "metrics_used_metric.json"), "w"))
    print("metrics cases:", len(cases))


def gen_cpp():
    from oracle import c_oracle as co
    co.build()
    assert co.ref_lib() is not None, "oracle/_ref/libref_eval.so missing"
    rng = np.random.default_rng(7)
    n_users, n_items, top_k = 40, 200, 20
    ratings = rng.normal(size=(n_users, n_items)).astype(np.float32)
    truth = [np.sort(rng.permutation(n_items)[: rng.integers(1, 30)]).astype(np.int32) for _ in range(n_users)]
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    indptr[1:] = np.cumsum([len(t) for t in truth])
    titems = np.concatenate(truth).astype(np.int32)
    metric = np.array([1, 2, 3, 4, 5], dtype=np.int32)   # precision, recall, map, ndcg, mrr (metric.h metric_dict)
    res = co.ref_evaluate_matrix(ratings.copy(), indptr, titems, metric, top_k)
    topk = co.ref_arg_top_k_2d(ratings.copy(), top_k)
    np.savez_compressed(os.path.join(HERE, "cpp_evaluator.npz"), ratings=ratings, truth_indptr=indptr,
                        truth_items=titems, metric=metric, top_k=top_k, results=res, arg_topk=topk)
    print("cpp evaluator golden:", res.shape, topk.shape)


def gen_douban():
    z = zipfile.ZipFile(os.path.join(REF, "data/douban/douban.zip"))
    pop = np.loadtxt(io.BytesIO(z.read("item_pop_seq_ori2.txt")))
    n_item = pop.shape[0]
    counts = np.zeros((10, n_item), dtype=np.int64)
    for t in range(10):
        for line in z.read(f"t_{t}.txt").decode().strip().split("\n"):
            parts = line.split()
            counts[t, int(parts[0])] = len(parts) - 1
    order = np.argsort(pop[:, 0])
    assert np.array_equal(pop[order, 0].astype(np.int64), np.arange(n_item))
    np.savez_compressed(os.path.join(HERE, "douban_pop_slice.npz"), counts=counts.astype(np.int32),
                        pop=pop[order, 1:])
    print("douban pop table:", pop.shape, "stage totals", counts.sum(axis=1))


def gen_oracle_step_eval():
    from oracle import pda_oracle as po
    rng = np.random.default_rng(424242)
    n_users, n_items, d, B, T, K = 256, 512, 64, 128, 9, 50
    U0 = po.xavier_init(n_users, d, 2021, 0)
    I0 = po.xavier_init(n_items, d, 2021, 1)
    # a little structure so that scores are not all ~0
    U0 = (U0 + rng.normal(0, 0.3, U0.shape)).astype(np.float32)
    I0 = (I0 + rng.normal(0, 0.3, I0.shape)).astype(np.float32)
    pop = rng.random((n_items, T + 1)) ** 2
    pop[rng.random(pop.shape) < 0.15] = 0.0
    P = po.train_pop_matrix(pop, 0.22)
    last, lin = po.eval_pops(pop, 0.22)
    om = po.OracleModel(n_users, n_items, d, 1e-2, 1e-3, B, "s_condition", U=U0, I=I0)
    batches, losses = [], []
    for s in range(3):
        users = rng.permutation(n_users)[:B].astype(np.int32)          # distinct users (rd.sample)
        perm = rng.permutation(n_items)
        pos, neg = perm[:B].astype(np.int32), perm[B:2 * B].astype(np.int32)   # distinct items: fp32 sums have one order
        t = rng.integers(0, T, B)
        pp, pn = P[pos, t].astype(np.float32), P[neg, t].astype(np.float32)
        batches.append((users, pos, neg, pp, pn))
        losses.append(om.train_step(users, pos, neg, pp, pn))
    uid = np.repeat(np.arange(n_users), 12)
    iid = rng.integers(0, n_items, len(uid))
    key = np.unique(uid.astype(np.int64) * n_items + iid)
    mptr, mitems, _ = po.build_csr(n_users, key // n_items, key % n_items)
    eval_users = rng.permutation(n_users)[:200].astype(np.int32)
    out = dict(U0=U0, I0=I0, U3=om.U, I3=om.I, losses=np.asarray(losses, dtype=np.float32), mask_indptr=mptr, mask_items=mitems,
               eval_users=eval_users, pop_last=last.astype(np.float32), pop_linear=lin.astype(np.float32), K=K)
    for s, b in enumerate(batches):
        for name, arr in zip(("users", "pos", "neg", "pos_pop", "neg_pop"), b):
            out[f"b{s}_{name}"] = arr
    for tag, rec, p in (("main", "main_branch", None), ("pda_last", "condition", last), ("pda_linear", "condition", lin)):
        ids, sc = po.recommend(om.U, om.I, eval_users, rec, K, mptr, mitems, pop=None if p is None else p.astype(np.float32),
                               return_scores=True)
        out[f"ids_{tag}"], out[f"scores_{tag}"] = ids.astype(np.int32), sc.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "oracle_step_eval.npz"), **out)
    print("oracle step/eval golden:", out["losses"])


if __name__ == "__main__":
    if os.path.exists(REF):
        gen_metrics()
        gen_cpp()
        gen_douban()
    gen_oracle_step_eval()
