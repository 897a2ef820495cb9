"""Pins the oracle against outputs of the REFERENCE's own runnable code (no GPU, no /root/reference needed):
the committed fixtures of tests/golden/ were produced by tests/golden/make_golden.py from MF/used_metric.py
(imported unmodified), the C++ evaluator headers (compiled to oracle/_ref) and the shipped Douban files."""
import json
import os

import numpy as np
import pytest

from oracle import pda_oracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def metric_cases():
    return json.load(open(os.path.join(GOLD, "metrics_used_metric.json")))


def test_get_performance_matches_used_metric(metric_cases):
    """oracle get_performance == MF/used_metric.py:69-80 on 61 recorded cases (abs 1e-12)."""
    assert len(metric_cases) >= 60
    for c in metric_cases:
        got = po.get_performance(c["truth"], c["ids"], c["Ks"])
        for k in ("recall", "precision", "ndcg", "hit_ratio"):
            assert np.allclose(got[k], c[k], rtol=0, atol=1e-12), (k, c["Ks"])


def test_survey_known_answer(metric_cases):
    """get_performance([1,5],[5,2,3,1],[2,4]) -> recall [0.5,1], ndcg [0.6131,0.8772]."""
    c = metric_cases[-1]
    assert c["truth"] == [1, 5] and c["ids"] == [5, 2, 3, 1]
    got = po.get_performance([1, 5], [5, 2, 3, 1], [2, 4])
    assert np.allclose(got["recall"], [0.5, 1.0]) and np.allclose(got["ndcg"], [0.6131471927654584, 0.8772153153380493])


def test_c_metrics_sum_matches_used_metric(metric_cases, c_oracle):
    """the C restatement (what the GPU metrics kernel is compared with) against the same recorded cases."""
    by_ks = {}
    for c in metric_cases[:-1]:
        by_ks.setdefault(tuple(c["Ks"]), []).append(c)
    for Ks, cases in by_ks.items():
        ids = np.array([c["ids"] for c in cases], dtype=np.int32)
        indptr = np.zeros(len(cases) + 1, dtype=np.int64)
        indptr[1:] = np.cumsum([len(c["truth"]) for c in cases])
        titems = np.concatenate([c["truth"] for c in cases]).astype(np.int32)
        got = c_oracle.metrics_sum(ids, np.arange(len(cases)), indptr, titems, list(Ks))
        for k in ("recall", "precision", "ndcg", "hit_ratio"):
            want = np.sum([c[k] for c in cases], axis=0)
            assert np.allclose(got[k], want, rtol=1e-12, atol=1e-12), (k, Ks)


@pytest.fixture(scope="module")
def cpp_gold():
    return np.load(os.path.join(GOLD, "cpp_evaluator.npz"))


def test_topk_matches_reference_arg_topk(cpp_gold):
    """oracle top-k (tf.nn.top_k order) == util/cython/include/arg_topk.h:arg_top_k_2d on tie-free scores."""
    K = int(cpp_gold["top_k"])
    assert np.array_equal(po.topk_ids(cpp_gold["ratings"], K), cpp_gold["arg_topk"])


def test_metrics_match_reference_cpp_evaluator(cpp_gold, c_oracle):
    """evaluate.h:cpp_evaluate_matrix reports cumulative metrics for k = 1..top_k (float accumulators); at every
    k its precision / recall / ndcg columns must equal the oracle's @k values (the two metric
    stacks coincide at the final K)."""
    K = int(cpp_gold["top_k"])
    ratings, indptr, titems = cpp_gold["ratings"], cpp_gold["truth_indptr"], cpp_gold["truth_items"]
    res = cpp_gold["results"].reshape(ratings.shape[0], len(cpp_gold["metric"]), K)
    col = {int(m): i for i, m in enumerate(cpp_gold["metric"])}      # 1 precision, 2 recall, 4 ndcg (metric.h:112-117)
    ids = po.topk_ids(ratings, K)
    Ks = list(range(1, K + 1))
    for u in range(ratings.shape[0]):
        truth = titems[indptr[u]:indptr[u + 1]]
        got = po.get_performance(truth, ids[u], Ks)
        assert np.allclose(got["precision"], res[u, col[1]], atol=2e-6)
        assert np.allclose(got["recall"], res[u, col[2]], atol=2e-6)
        assert np.allclose(got["ndcg"], res[u, col[4]], atol=2e-6)
    s = c_oracle.metrics_sum(ids, np.arange(ratings.shape[0]), indptr, titems, [5, K])
    assert np.allclose(s["recall"], res[:, col[2], [4, K - 1]].sum(0), atol=1e-4)
    assert np.allclose(s["ndcg"], res[:, col[4], [4, K - 1]].sum(0), atol=1e-4)


def test_reference_evaluator_live_when_built(cpp_gold, c_oracle):
    """When oracle/_ref/libref_eval.so travelled with the repo, re-run the reference code itself."""
    if c_oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    K = int(cpp_gold["top_k"])
    res = c_oracle.ref_evaluate_matrix(cpp_gold["ratings"].copy(), cpp_gold["truth_indptr"], cpp_gold["truth_items"],
                                       cpp_gold["metric"], K)
    assert np.array_equal(res, cpp_gold["results"])
    assert np.array_equal(c_oracle.ref_arg_top_k_2d(cpp_gold["ratings"].copy(), K), cpp_gold["arg_topk"])


def test_pop_table_formula_reproduces_shipped_douban_table():
    """pop_pre.py:12-42 restated (pop_table_from_stage_counts) on the per-stage counts of the shipped t_k.txt
    reproduces the shipped item_pop_seq_ori2.txt exactly (SURVEY section 4 known answer)."""
    z = np.load(os.path.join(GOLD, "douban_pop_slice.npz"))
    counts, pop = z["counts"], z["pop"]
    assert counts.shape == (10, 26047) and pop.shape == (26047, 10)
    assert counts.sum(1).tolist() == [973800, 795705, 705637, 636706, 631387, 524913, 733527, 838223, 786067, 548253]
    got = po.pop_table_from_stage_counts(counts)
    assert np.abs(got - pop).max() <= 1e-12
    # eval / train popularity preparation (train_new_api.py:952-959, 988-990)
    P = po.train_pop_matrix(pop, 0.22)
    assert P.shape == (26047, 9) and P.dtype == np.float32
    assert np.array_equal(P, np.power(pop[:, :-1], 0.22).astype(np.float32))
    last, lin = po.eval_pops(pop, 0.22)
    assert np.array_equal(last, np.power(pop[:, -2], 0.22).astype(np.float32))
    raw = pop[:, -2] + 0.5 * (pop[:, -2] - pop[:, -3])
    assert np.allclose(lin, np.power(np.clip(np.where(raw <= 0, 1e-9, raw), None, 1.0), 0.22).astype(np.float32))
    assert (P == 0).sum() > 1000          # exact zeros survive pop ** gamma


# ---------------------------------------------------------------------------------------------
# oracle-generated golden vectors for the step and the scoring: pins both oracles against drift
# ---------------------------------------------------------------------------------------------
def _step_eval_golden():
    return np.load(os.path.join(GOLD, "oracle_step_eval.npz"))


def test_numpy_and_c_oracle_reproduce_the_step_and_eval_golden(c_oracle):
    from oracle import pda_oracle as po
    g = _step_eval_golden()
    B = len(g["b0_users"])
    om = po.OracleModel(g["U0"].shape[0], g["I0"].shape[0], g["U0"].shape[1], 1e-2, 1e-3, B, "s_condition", U=g["U0"], I=g["I0"])
    cm = c_oracle.CModel(g["U0"], g["I0"], 1e-2, 1e-3, B, "s_condition")
    for s in range(3):
        b = [g[f"b{s}_{k}"] for k in ("users", "pos", "neg", "pos_pop", "neg_pop")]
        ln, lc = om.train_step(*b), cm.train_step(*b)
        assert np.array_equal(np.asarray(ln, np.float32).view(np.int32), g["losses"][s].view(np.int32)), s
        assert np.allclose(lc, g["losses"][s], rtol=1e-6)          # libm logf of the C build: 1e-6, everything else exact
    for tab, U, I in (("numpy", om.U, om.I), ("c", cm.U, cm.I)):
        assert np.array_equal(U.view(np.int32), g["U3"].view(np.int32)), tab
        assert np.array_equal(I.view(np.int32), g["I3"].view(np.int32)), tab
    K = int(g["K"])
    for tag, rec, p in (("main", "main_branch", None), ("pda_last", "condition", g["pop_last"]),
                        ("pda_linear", "condition", g["pop_linear"])):
        ids, sc = c_oracle.recommend(g["U3"], g["I3"], g["eval_users"], rec, K, g["mask_indptr"], g["mask_items"], pop=p)
        assert np.array_equal(ids, g[f"ids_{tag}"]), tag
        assert np.array_equal(sc.view(np.int32), g[f"scores_{tag}"].view(np.int32)), tag
