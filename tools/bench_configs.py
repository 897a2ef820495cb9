"""bench.py --config douban_pd | douban_pda_eval | kwai: BASELINE.json configs[1..3] as secondary bench lines.

  douban_pd        PD (s_condition) on the shipped Douban set, d=64, gamma=0.22, B=2048: steps/s and triples/s of one
                   epoch (3236 steps).  At B=2048 a step moves 3.2 MB: it is LAUNCH-bound, not HBM-bound;
                   the line says so and carries the per-kernel times.
  douban_pda_eval  PDA evaluation on Douban: valid (6847) and test (15974) users x 26047 items, top-50 + metrics.
  kwai             PD on a Kwai-shaped synthetic set (37663 x 128879, ~7M pairs, 9 stages, d=64, gamma=0.16):
                   train + eval at 1/2/4/8 GPUs with GLOBAL batch 2048 (torchrun for N > 1).
Each prints ONE JSON line with the CPU port (oracle/) timed beside it on rank 0.  Needs ./data/douban caches
(tools/stage_douban.py) for the Douban configs.
"""
from __future__ import annotations

import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _douban():
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


def _quiet_stdout():
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return out


def douban_pd(a):
    import pda_b200
    from oracle import c_oracle as co
    from oracle import pda_oracle as po
    out = _quiet_stdout()
    d, pop = _douban()
    gamma, B, dim = 0.22, 2048, 64
    P = po.train_pop_matrix(pop, gamma)
    m = pda_b200.PDAModel(d.n_users, d.n_items, dim, train="s_condition", batch_size=B, lr=1e-2, regs=1e-3, seed=2021)
    m.set_train_csr(d.train_indptr, d.train_items, d.train_times, unique_times=d.unique_times)
    m.set_train_pop(P)
    n_batch = d.n_train // B + 1
    m.train_sampled(2020, 0, 0, 200, B)
    m.synchronize()
    m.profile(True)
    t0 = time.perf_counter()
    m.train_sampled(2020, 1, 0, n_batch, B)
    m.synchronize()
    dt = time.perf_counter() - t0
    pr = m.profile_read()
    m.profile(False)
    # the sess.run-shaped host API, one call per batch
    b = [m.sample_batch(2020, 2, s, B) for s in range(64)]
    m.train_step(b[0]["users"], b[0]["pos"], b[0]["neg"], b[0]["pos_pop"], b[0]["neg_pop"])
    t0 = time.perf_counter()
    for x in b:
        m.train_step(x["users"], x["pos"], x["neg"], x["pos_pop"], x["neg_pop"])
    dt_host = (time.perf_counter() - t0) / len(b)
    # CPU port on the same batches
    co.build()
    cores = co.set_num_threads(os.cpu_count() or 1)
    ref = co.CModel(po.xavier_init(d.n_users, dim, 2021, 0), po.xavier_init(d.n_items, dim, 2021, 1), 1e-2, 1e-3, B, "s_condition")
    active = np.nonzero(np.diff(d.train_indptr) > 0)[0]
    n_cpu = 200
    t0 = time.perf_counter()
    for s in range(n_cpu):
        x = co.sample_batch(2020, 1, s, B, active, d.train_indptr, d.train_items, d.train_times, d.n_items,
                            np.asarray(d.unique_times), P)
        ref.train_step(x["users"], x["pos"], x["neg"], x["pos_pop"], x["neg_pop"])
    dt_cpu = (time.perf_counter() - t0) / n_cpu
    step_ms, step_n = pr["bpr_step"]
    adam_ms, adam_n = pr["adam"]
    line = {"config": {"workload": "Douban PD (s_condition) d=64 gamma=0.22 B=2048, one epoch = %d steps" % n_batch},
            "metric": "bpr_triples_per_sec", "value": n_batch * B / dt, "unit": "triples/s", "steps_per_sec": n_batch / dt,
            "ms_per_step": dt / n_batch * 1e3, "n_gpus": 1, "dtype": "f32", "data": "douban (shipped with the reference)",
            "bound": "launch (3.2 MB per step: 4 kernels of 3-15 us each; HBM fraction is not meaningful at B=2048, SURVEY 7)",
            "kernels": {"bpr_step_us": step_ms / max(step_n, 1) * 1e3, "adam_us": adam_ms / max(adam_n, 1) * 1e3,
                        "sampler_us": pr["sampler"][0] / max(pr["sampler"][1], 1) * 1e3},
            "e2e": {"value": B / dt_host, "unit": "triples/s", "api": "PDAModel.train_step (one sess.run-shaped call per host batch)",
                    "h2d_bytes_per_step": 20 * B, "d2h_bytes_per_step": 12},
            "cpu_baseline": {"value": B / dt_cpu, "unit": "triples/s", "cores": cores, "kind": "port",
                             "sample": "%d steps of the C/OpenMP port (sampler + step + dense Adam)" % n_cpu}}
    out.write(json.dumps(line) + "\n")
    m.close()


def douban_pda_eval(a):
    import pda_b200
    from oracle import c_oracle as co
    from oracle import pda_oracle as po
    from pda_b200.evaluation import evaluation
    out = _quiet_stdout()
    d, pop = _douban()
    gamma, B, dim = 0.22, 2048, 64
    P = po.train_pop_matrix(pop, gamma)
    last, lin = po.eval_pops(pop, gamma)
    m = pda_b200.PDAModel(d.n_users, d.n_items, dim, train="s_condition", batch_size=B, lr=1e-2, regs=1e-3, seed=2021)
    m.set_train_csr(d.train_indptr, d.train_items, d.train_times, unique_times=d.unique_times)
    m.set_train_pop(P)
    m.train_sampled(2020, 0, 0, 2 * (d.n_train // B + 1), B)        # two epochs: scores spread beyond the init
    m.synchronize()
    U, I = m.get_table("user_embedding"), m.get_table("item_embedding")
    co.build()
    cores = co.set_num_threads(os.cpu_count() or 1)
    IT = np.ascontiguousarray(I.T)
    res = {}
    for who, ul, ptr, items in (("valid", d.valid_user_list, d.valid_indptr, d.valid_items),
                                ("test", d.test_user_list, d.test_indptr, d.test_items)):
        users = np.asarray(list(ul.keys()), dtype=np.int32)
        m.do_recommendation(users[:256], None, "condition", pos_pop=last, K=50)
        m.profile(True)
        t0 = time.perf_counter()
        ids = m.do_recommendation(users, None, "condition", pos_pop=last, K=50)
        s = m.metrics_sum(ids, users, ptr, items, [20, 50])
        wall = time.perf_counter() - t0
        pr = m.profile_read()
        m.profile(False)
        kms = pr["eval_exact"][0] + pr["eval_tensor"][0]
        t0 = time.perf_counter()
        rid, _ = co.recommend(U, None, users, "condition", 50, d.train_indptr, d.train_items, pop=last, IT=IT)
        cs = co.metrics_sum(rid, users, ptr, items, [20, 50])
        dt_cpu = time.perf_counter() - t0
        pairs = len(users) * d.n_items
        res[who] = {"users": int(len(users)), "pairs": int(pairs), "kernel_ms": kms, "pairs_per_sec_kernels": pairs / (kms * 1e-3),
                    "pairs_per_sec_e2e": pairs / wall, "tflops": pairs * 2 * dim / (kms * 1e-3) / 1e12,
                    "filter_stats": m.tc_last_stats() if pr["eval_tensor"][1] else None,
                    "recall@20": float(s["recall"][0] / len(users)), "ndcg@20": float(s["ndcg"][0] / len(users)),
                    "ids_equal_oracle": bool(np.array_equal(ids, rid)),
                    "recall@20_oracle": float(cs["recall"][0] / len(users)),
                    "cpu_baseline": {"value": pairs / dt_cpu, "unit": "pairs/s", "cores": cores, "kind": "port",
                                     "sample": "the full eval (all %d users): scoring + mask + top-50 + metrics" % len(users)}}
    line = {"config": {"workload": "Douban PDA eval (condition, last-stage pop^0.22): U.I^T + pop + train mask + top-50 + Recall/NDCG"},
            "metric": "eval_user_item_pairs_per_sec", "value": res["test"]["pairs_per_sec_e2e"], "unit": "pairs/s", "n_gpus": 1,
            "dtype": "bf16 filter + f32 exact rescoring", "data": "douban (shipped with the reference)",
            "bound": "launch / latency (0.2-0.5 ms of kernels per eval: nine launches of 5-120 us)", "valid": res["valid"], "test": res["test"]}
    out.write(json.dumps(line) + "\n")
    m.close()


def kwai_interactions(n_users, n_items, mean_deg, n_stages, seed, empty_frac=0.002):
    """Kwai-shaped synthetic interactions (Kwai is not shipped): unique (user, item) pairs, Zipf item draw, uniform stage"""
    rng = np.random.default_rng(seed)
    deg = 1 + rng.poisson(mean_deg, n_users)
    deg[rng.random(n_users) < empty_frac] = 0
    uid = np.repeat(np.arange(n_users), deg)
    w = 1.0 / (1.0 + np.arange(n_items))
    iid = rng.choice(n_items, size=len(uid), p=w / w.sum())
    key = np.unique(uid.astype(np.int64) * n_items + iid)
    uid, iid = key // n_items, key % n_items
    return uid, iid, rng.integers(0, n_stages, len(uid))


def kwai(a):
    import torch
    import torch.distributed as dist
    import pda_b200
    from oracle import pda_oracle as po
    from pda_b200.parallel import ShardedEvaluator, ShardedTrainer, shard_range
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    out = _quiet_stdout()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    NU, NI, T, dim, gamma, Bg = 37663, 128879, 9, 64, 0.16, 2048
    B = Bg // world
    uid, iid, t = kwai_interactions(NU, NI, 250, T, seed=11)
    rng = np.random.default_rng(12)
    pop = rng.random((NI, T + 1)) ** 3
    pop[rng.random(pop.shape) < 0.2] = 0.0
    pop /= pop.max(axis=0, keepdims=True)
    # held-out pairs of the eval stage: 20 random items per user
    truth = rng.integers(0, NI, (NU, 20))
    lo, hi = shard_range(NU, world, rank)
    sel = (uid >= lo) & (uid < hi)
    indptr, items, times = po.build_csr(hi - lo, uid[sel] - lo, iid[sel], t[sel])
    P = po.train_pop_matrix(pop, gamma)
    last, _ = po.eval_pops(pop, gamma)
    m = pda_b200.PDAModel(hi - lo, NI, dim, train="s_condition", batch_size=Bg, lr=1e-2, regs=1e-3, device=local, max_batch=B, seed=2021)
    m.set_train_csr(indptr, items, times, unique_times=np.arange(T))
    m.set_train_pop(P)
    tr = ShardedTrainer(m, world, rank)
    stream = torch.cuda.current_stream().cuda_stream
    n_steps = max(a.steps, 200)
    tr.train_sampled(2020 + rank, 0, 0, 50, B, stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tr.train_sampled(2020 + rank, 1, 0, n_steps, B, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    # eval: every rank its active users x all items
    users = np.nonzero(np.diff(indptr) > 0)[0].astype(np.int32)
    tptr = np.arange(0, (hi - lo + 1) * 20, 20, dtype=np.int64)
    titems = np.sort(truth[lo:hi], axis=1).reshape(-1).astype(np.int32)
    ev = ShardedEvaluator(m, world, rank)
    ev.eval(users[:256], tptr, titems, [20, 50], rec_type="condition", pos_pop=last, K=50, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = ev.eval(users, tptr, titems, [20, 50], rec_type="condition", pos_pop=last, K=50, device="cuda")
    torch.cuda.synchronize()
    tev = torch.tensor([time.perf_counter() - t0, float(len(users))], device="cuda", dtype=torch.float64)
    if world > 1:
        mx = tev.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tev, op=dist.ReduceOp.SUM)
        t_eval, n_eval = float(mx[0].item()), float(tev[1].item())
    else:
        t_eval, n_eval = float(tev[0].item()), float(tev[1].item())
    if rank == 0:
        line = {"config": {"workload": "Kwai-shaped synthetic PD: 37663 users x 128879 items, %d train pairs, d=64, gamma=0.16, "
                                       "GLOBAL batch 2048 (%d per GPU)" % (len(uid), B)},
                "metric": "bpr_triples_per_sec", "value": n_steps * Bg / (ms * 1e-3), "unit": "triples/s", "n_gpus": world,
                "steps": n_steps, "ms_per_step": ms / n_steps, "scaling": "strong (global batch fixed at 2048 for parity runs, SURVEY 8e)",
                "bound": "launch / latency: a 2048-triple step is 3 MB of traffic",
                "eval": {"value": n_eval * NI / t_eval, "unit": "pairs/s", "users": int(n_eval), "seconds": t_eval,
                         "recall@20": float(r["recall"][0])}, "dtype": "f32", "data": "synthetic (Kwai-shaped)"}
        out.write(json.dumps(line) + "\n")
    m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main(a):
    sys.path.insert(0, ROOT)
    {"douban_pd": douban_pd, "douban_pda_eval": douban_pda_eval, "kwai": kwai}[a.config](a)
