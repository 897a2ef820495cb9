"""Times the two eval back ends (exact CUDA-core kernel vs tcgen05 filter) on a Douban-shaped and a synthetic-shaped
problem with random tables; prints pairs/s, TFLOP/s (2*d per pair) and the filter statistics.  GPU only."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pda_b200  # noqa: E402


def run(n_users, n_items, d, M, rec_type, deg, reps=5, backends=("exact", "tensor"), scale=1.0):
    rng = np.random.default_rng(0)
    m = pda_b200.PDAModel(n_users, n_items, d, train="s_condition", batch_size=64, seed=2021)
    if scale != 1.0:
        m.set_table("user_embedding", m.get_table("user_embedding") * scale)
        m.set_table("item_embedding", m.get_table("item_embedding") * scale)
    degs = np.minimum(1 + rng.poisson(deg, n_users), n_items // 4)
    if n_users <= 200000:
        rows = [np.sort(rng.choice(n_items, k, replace=False)) for k in degs]
    else:
        draw = np.sort(rng.integers(0, n_items, (n_users, int(degs.max()))), axis=1)
        rows = [np.unique(draw[u, :degs[u]]) for u in range(n_users)]
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    indptr[1:] = np.cumsum([len(r) for r in rows])
    m.set_train_csr(indptr, np.concatenate(rows).astype(np.int32))
    pop = (rng.random(n_items) ** 0.22).astype(np.float32)
    users = np.arange(M, dtype=np.int32)
    out = {}
    ref = None
    for be in backends:
        m.do_recommendation(users[:1024], None, rec_type, pos_pop=pop, K=50, backend=be)
        m.profile(True)
        t0 = time.perf_counter()
        for _ in range(reps):
            ids = m.do_recommendation(users, None, rec_type, pos_pop=pop, K=50, backend=be)
        wall = (time.perf_counter() - t0) / reps
        pr = m.profile_read()
        m.profile(False)
        kms = (pr["eval_exact"][0] + pr["eval_tensor"][0]) / reps
        pairs = M * n_items
        out[be] = dict(kernel_ms=kms, wall_ms=wall * 1e3, gpairs_s=pairs / kms / 1e6, tflops=pairs * 2 * d / kms / 1e9)
        if be == "tensor":
            out[be]["stats"] = m.tc_last_stats()
        if ref is None:
            ref = ids
        else:
            out[be]["ids_equal_exact"] = bool(np.array_equal(ref, ids))
    m.close()
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "douban"
    bes = tuple(sys.argv[2].split(",")) if len(sys.argv) > 2 else ("exact", "tensor")
    if which == "douban":
        for rt in ("main_branch", "condition"):
            print(json.dumps({"shape": "douban-like 47890x26047 d=64 M=15974", "rec_type": rt,
                              **run(47890, 26047, 64, 15974, rt, 138, backends=bes)}))
    else:
        for rt in ("condition",):
            print(json.dumps({"shape": "synthetic 65536x1000000 d=128 M=16384", "rec_type": rt,
                              **run(65536, 1000000, 128, 16384, rt, 24, reps=2, backends=bes)}))
