"""The NeuRec native-evaluator entry points over the GPU (row b4): pda_arg_top_k_2d_host / pda_evaluate_matrix_host against
the REFERENCE's own C++ (evaluator/backend/cpp/include/evaluate.h, util/cython/include/arg_topk.h): the committed golden
vectors recorded from it, and the live library (oracle/_ref) on all five metrics when it travelled with the repo."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def neurec():
    import pda_b200
    assert pda_b200.load().pda_device_count() >= 1
    from pda_b200 import neurec
    return neurec


def test_arg_topk_and_evaluate_matrix_reproduce_the_reference_golden(neurec):
    g = np.load(os.path.join(GOLD, "cpp_evaluator.npz"))
    K = int(g["top_k"])
    assert np.array_equal(neurec.arg_topk(g["ratings"], K), g["arg_topk"])
    truth = [g["truth_items"][g["truth_indptr"][u]:g["truth_indptr"][u + 1]] for u in range(g["ratings"].shape[0])]
    res = neurec.apk_evaluate_matrix(g["ratings"], truth, [int(m) for m in g["metric"]], top_k=K)
    assert res.shape == g["results"].shape
    assert np.allclose(res, g["results"], rtol=0, atol=1e-6)


def test_all_five_metrics_against_the_live_reference_evaluator(neurec, c_oracle):
    if c_oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    rng = np.random.default_rng(11)
    n_users, n_items, K = 257, 3001, 50
    ratings = rng.permutation(n_users * n_items).reshape(n_users, n_items).astype(np.float32)      # tie-free
    ratings[:, ::7] = -np.inf                                                                      # masked train items
    ratings += rng.random((n_users, 1)).astype(np.float32)
    truth = [rng.choice(n_items, rng.integers(1, 60), replace=False) for _ in range(n_users)]
    ptrs = np.zeros(n_users + 1, dtype=np.int64)
    ptrs[1:] = np.cumsum([len(t) for t in truth])
    want = c_oracle.ref_evaluate_matrix(ratings.copy(), ptrs, np.concatenate(truth).astype(np.int32), np.array([1, 2, 3, 4, 5]), K)
    got = neurec.apk_evaluate_matrix(ratings, truth, ["Precision", "Recall", "MAP", "NDCG", "MRR"], top_k=K)
    assert np.allclose(got, want, rtol=0, atol=2e-6)
    finite = np.where(np.isfinite(ratings), ratings, -3.0e38)
    assert np.array_equal(neurec.arg_topk(finite, K), c_oracle.ref_arg_top_k_2d(finite.copy(), K))
