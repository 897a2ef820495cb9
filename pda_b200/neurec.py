"""NeuRec-style evaluator front (the reference's optional native backend) over the GPU library.

Mirrors util/cython/arg_topk.pyx:16-35 (`arg_topk`) and evaluator/backend/cpp/cpp_evaluator.pyx:15-42
(`apk_evaluate_matrix`): same arguments, same result shapes; `thread_num` is accepted and ignored (the GPU does the work)."""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check, ptr

METRICS = {"Precision": 1, "Recall": 2, "MAP": 3, "NDCG": 4, "MRR": 5}


def arg_topk(matrix, topk=50, thread_num=1):
    m = np.ascontiguousarray(matrix, dtype=np.float32)
    rows, cols = m.shape
    out = np.empty((rows, topk), dtype=np.int32)
    check(_lib.load().pda_arg_top_k_2d_host(ptr(m), cols, rows, topk, ptr(out)))
    return out


def apk_evaluate_matrix(ratings, test_items, metric, top_k=50, thread_num=1):
    """ratings: [n_users, n_items] float32; test_items: list of iterables of item ids (one per user);
    metric: list of names or ids -> float32 [n_users, len(metric) * top_k]"""
    r = np.ascontiguousarray(ratings, dtype=np.float32)
    n_users, n_items = r.shape
    ptrs = np.zeros(n_users + 1, dtype=np.int64)
    rows = [np.asarray(sorted(t), dtype=np.int32) for t in test_items]
    ptrs[1:] = np.cumsum([len(x) for x in rows])
    items = np.concatenate(rows).astype(np.int32) if len(rows) and ptrs[-1] > 0 else np.zeros(1, dtype=np.int32)
    ms = np.asarray([METRICS[x] if isinstance(x, str) else int(x) for x in metric], dtype=np.int32)
    out = np.empty((n_users, len(ms) * top_k), dtype=np.float32)
    check(_lib.load().pda_evaluate_matrix_host(ptr(r), n_items, n_users, ptr(ptrs), ptr(items), ptr(ms), len(ms), top_k, ptr(out)))
    return out
