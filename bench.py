#!/usr/bin/env python
"""bench.py -- headline benchmark of the PDA hot path on B200 (contract: see DESIGN.md section 7).

  python bench.py [--gpus N] [--steps K] [--warmup W]                    # this repo's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # CPU reference arm (oracle port)
  python bench.py --config douban_pd | douban_pda_eval | kwai            # BASELINE configs[1..3] (secondary lines)

A "step" = one pass of the hot path over one batch: device sampler -> fused BPR step kernel
(gather, dots, ELU'/pop^gamma, log-sigmoid BPR loss + L2, gradient scatter-add, Adam of the user rows) ->
TF1-semantics Adam sweep of the item table.  Default workload (BASELINE.json configs[4], fits one B200):
synthetic 10M users x 1M items, d=128, PD (--train s_condition), gamma=0.16.  `value` = triples/s with
everything resident in HBM; `e2e` = the same metric through the host API (pinned host batches -> H2D ->
step -> D2H of the 3 loss scalars, every step).  The line also carries `parity`: the CUDA path checked
against the CPU oracle ON THIS WORKLOAD (outside the timed regions).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GAMMA = 0.16
SEED_DATA, SEED_SAMPLER, SEED_INIT = 2020, 2020, 2021
PARITY_STEPS = 5


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="synthetic", choices=["synthetic", "douban_pd", "douban_pda_eval", "kwai"])
    ap.add_argument("--users", type=int, default=10_000_000)
    ap.add_argument("--items", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=1 << 20, help="triples per step per GPU")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--regs", type=float, default=1e-3)
    ap.add_argument("--eval-users", type=int, default=65536, help="users scored against all items (blocks of 32768)")
    ap.add_argument("--eval-tables", default="both", choices=["trained", "fitted", "both"],
                    help="trained = the tables after this run's few Adam steps; fitted = N(0,1)/sqrt(d) tables with skewed "
                         "item norms and pop^6 (a fitted model's candidate load)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU legs (cpu_baseline + parity)")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=8, help="timed steps of the cpu_baseline leg (about 0.4 s each on 16 cores)")
    ap.add_argument("--adam", default="auto", choices=["auto", "lazy", "lazy_users", "dense"],
                    help="how the TF1 every-row Adam sweep is evaluated (bit-identical results; see DESIGN.md 5.2)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=float(j["hbm_gbs"]), bf16=float(j["bf16_tflops"]), bf16_sus=float(j["bf16_tflops_sustained"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples DURING the timed region (profiling recipe's clocks line).  Started well
    before the region (nvidia-smi needs a few hundred ms to come up); summary(t0, t1) keeps the samples inside it."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0=None, t1=None):
        ok = [(t, r) for t, r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        inside = [(t, r) for t, r in ok if t0 is not None and t0 <= t <= t1 + 0.03]
        scope = "timed region"
        if not inside:      # region shorter than the sampling period: the samples under load around it
            inside, scope = [(t, r) for t, r in ok if t0 is None or t0 - 0.5 <= t <= t1 + 0.1], "timed region +- 0.5 s (under load)"
        sm = [float(r[0]) for _, r in inside]
        mx = [float(r[1]) for _, r in inside if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for _, r in inside for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "scope": scope}


def wrap_ptr(t):
    return int(t.data_ptr())


def ncu_traffic():
    """DRAM bytes per launch of the kernels of THIS workload from the committed ncu --set full captures (profiles/)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            j = json.load(open(os.path.join(ROOT, "profiles", name)))
            j["_file"] = "profiles/" + name
            return j
        except (OSError, ValueError):
            continue
    return {}


def make_config(a, world):
    """identical for our arm and the reference arm (the driver compares the two lines)"""
    users_local = a.users // world
    return {"workload": f"synthetic {a.users} users x {a.items} items d={a.dim}, PD (s_condition) gamma={GAMMA}, "
                        f"TF1 every-row Adam semantics, B={a.batch} triples/step/GPU",
            "users": a.users, "items": a.items, "d": a.dim, "batch_per_gpu": a.batch, "global_batch": a.batch * world,
            "parallelism": f"user-shard x{world}, items replicated" + (" + item-gradient exchange over NVLink (fused multicast kernel from 4 ranks on, NCCL reduce-scatter / all-gather below)" if world > 1 else ""),
            "l2_policy": "tables >> L2 (user table %.1f GB per rank): no flush needed" % (users_local * a.dim * 4 / 1e9)}


def dev_view(ptr, shape, dev, typestr="<f4"):
    import torch
    from pda_b200.parallel import _DevArray
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=dev)


# ------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    import pda_b200
    from pda_b200 import synth
    from pda_b200.parallel import ShardedTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        a.gpus = world
    torch.cuda.set_device(local)
    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 from here on (NCCL's version banner
    # under NCCL_DEBUG=VERSION/WARN, library chatter) goes to stderr; the line itself is written to the saved fd
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        # high-priority NCCL stream: the exchange must get its CTAs although full-occupancy kernels of the rank are runnable
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    dev = torch.device("cuda", local)
    pk = peaks()
    B, d = a.batch, a.dim
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()

    # users are partitioned over ranks (rank-local sampling, no user traffic); items replicated
    users_local = a.users // world
    ds = synth.make_synthetic(users_local, a.items, seed=SEED_DATA + rank, device=dev)
    P = synth.train_pop_matrix_torch(ds["pop"], GAMMA)
    model = pda_b200.PDAModel(users_local, a.items, d, train="s_condition", batch_size=B * world, lr=a.lr, regs=a.regs,
                              device=local, max_batch=B, seed=SEED_INIT, init=True)
    model.set_train_csr_device(wrap_ptr(ds["indptr"]), wrap_ptr(ds["items"]), wrap_ptr(ds["times"]), ds["nnz"],
                               wrap_ptr(ds["active"]), ds["active"].numel(), unique_times=np.arange(ds["n_stages"] - 1))
    model.set_train_pop(P.cpu().numpy())
    if a.adam != "auto":
        model.set_adam_mode(a.adam)
    else:
        a.adam = "lazy_users"    # what the library picks here: 10M user rows vs 2^20 refs/step -> lazy; 1M item rows vs 2^21 -> dense
    trainer = ShardedTrainer(model, world, rank)      # world > 1: lazy user table + dense (exchanged) item table
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run_steps(step0, n):
        trainer.train_sampled(SEED_SAMPLER + rank, 0, step0, n, B, stream)

    # ---- parity prologue (untimed): the first steps one by one, losses and the step-0 batch kept for the CPU oracle ----
    gpu_first = {"losses": []}
    b0 = model.sample_batch(SEED_SAMPLER + rank, 0, 0, B)
    gpu_first["step0"] = {k: b0[k].copy() for k in ("users", "pos", "neg", "pos_pop", "neg_pop")}
    for s in range(PARITY_STEPS):
        run_steps(s, 1)
        gpu_first["losses"].append([float(x) for x in model.read_loss(stream)])
    step_no = PARITY_STEPS
    # ---- replay steady state: a sampled user row has skipped ~U/B steps; reach that lag distribution before timing ----
    age = max(0, int(math.ceil(3.0 * users_local / B)) - PARITY_STEPS) if a.adam != "dense" else 0
    run_steps(step_no, age) if age else None
    step_no += age

    # ---- device-resident run: `value` ----
    run_steps(step_no, a.warmup)
    step_no += a.warmup
    barrier()
    model.profile(True)
    if a.adam != "dense":
        model.adam_stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        trainer.profile(True)       # per-phase events of the gradient exchange (rank-local, on the compute stream)
    barrier()
    tw0 = time.time()
    e0.record()
    run_steps(step_no, a.steps)
    e1.record()
    barrier()
    tw1 = time.time()
    step_no += a.steps
    ms = e0.elapsed_time(e1)
    exch = trainer.profile_summary() if world > 1 else None
    if world > 1:
        trainer.profile(False)
    prof = model.profile_read()
    model.profile(False)
    rows_updated, row_steps_replayed = model.adam_stats(reset=True) if a.adam != "dense" else (0, 0)
    loss = model.read_loss(stream)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = a.steps * B * world / (ms * 1e-3)
    clk = clocks.summary(tw0, tw1) if rank == 0 else None

    # ---- end-to-end through the host API: `e2e` ----
    e2e = None
    if not a.no_e2e:
        n_e2e = min(a.steps, 10)
        keys = ("users", "pos", "neg", "pos_pop", "neg_pop")
        # host batches live in pinned memory (cudaHostAlloc through the C ABI), [n, B] per array
        pin = {k: model.pinned_array((n_e2e + 2, B), np.float32 if k.endswith("pop") else np.int32) for k in keys}
        for s in range(n_e2e + 2):
            b = model.sample_batch(SEED_SAMPLER + rank, 1, s, B)
            for k in keys:
                pin[k][s] = b[k]
        if world == 1:
            fn = lambda lo, hi: model.train_steps(*(pin[k][lo:hi] for k in keys))
            api = "PDAModel.train_steps (pda_train_steps_host: n pinned host batches, copies pipelined with the steps)"
        else:
            fn = lambda lo, hi: trainer.train_steps_host(*(pin[k][lo:hi] for k in keys), stream=stream)
            api = ("ShardedTrainer.train_steps_host (pda_stage_batch_host_async on a copy stream under the previous step's "
                   "NCCL exchange + split step)")
        fn(0, 2)
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        l3 = fn(2, n_e2e + 2)[-1]
        t1.record()
        barrier()
        wall = time.perf_counter() - w0
        ems = max(t0.elapsed_time(t1), 0.0)
        te = torch.tensor([max(ems * 1e-3, wall)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": n_e2e * B * world / float(te.item()), "unit": "triples/s", "h2d_bytes_per_step": 20 * B,
               "d2h_bytes_per_step": 12, "steps": n_e2e, "api": api, "last_loss": [float(x) for x in l3]}

    # ---- multi-GPU consistency (untimed): replicas identical, and the sharded step == one process on the union batch ----
    par_multi = multi_gpu_parity(world, rank, local, dev, model, trainer) if world > 1 else None

    # ---- eval: all-items scoring + pop adjust + mask + top-50 (pairs/s) ----
    ev, ev_cpu_inputs = None, None
    if not a.no_eval:
        ev, ev_cpu_inputs = run_eval(a, model, ds, dev, world, rank, pk)

    # ---- per-kernel device times (CUDA events on the launching stream) and rooflines ----
    step_ms, step_n = prof["bpr_step"]
    adam_ms, adam_n = prof["adam"]
    cat_ms, cat_n = prof["adam_catchup"]
    samp_ms, samp_n = prof["sampler"]
    # SURVEY 8d counts 3 rows gathered + 3 gradient rows written + ids/pops = 24d+20 B per triple.  With the user row's
    # Adam update fused into the kernel (lazy user table + distinct users) the user gradient never reaches HBM and the
    # row's W read is the gather itself: what the kernel moves per triple is p, n gathered (8d) + two item-gradient rows
    # reduced (8d) + W, m, v of the user row read and written (24d) + ids/pops (20) = 40d+20 B.
    fused = a.adam in ("lazy", "lazy_users") and os.environ.get("PDA_FUSE_USER_ADAM", "1") != "0"
    bytes_triple = (40 * d + 20) if fused else (24 * d + 20)
    step_bytes = bytes_triple * B
    per_launch = step_ms / max(step_n, 1)
    step_gbs = step_bytes / (per_launch * 1e-3) / 1e9 if step_ms > 0 else 0.0
    # rows the Adam kernels outside the step kernel handle per step: dense-swept tables + lazily updated rows
    adam_rows = rows_updated / max(a.steps, 1)
    if a.adam == "dense":
        adam_rows += users_local + a.items
    elif world > 1:
        adam_rows += a.items / world          # every rank sweeps its row slice
    elif a.adam == "lazy_users":
        adam_rows += a.items
    adam_bytes = adam_rows * d * 4 * 6                                       # W, m, v read + written
    adam_gbs = adam_bytes / (adam_ms / max(a.steps, 1) * 1e-3) / 1e9 if adam_ms > 0 else 0.0
    kern = {
        "bpr_step": {"ms_per_launch": per_launch, "share_of_step": step_ms / ms, "achieved": step_gbs,
                     "frac": step_gbs / pk["hbm"], "algorithmic_bytes": step_bytes, "unit": "GB/s",
                     "bytes_model": "B*(40d+20): p,n gathered + 2 item-gradient rows reduced + W,m,v of the user row read+written "
                                    "(fused Adam) + ids/pops" if fused else "B*(24d+20) (SURVEY 8d)",
                     "bpr_only_gbs": (24 * d + 20) * B / (per_launch * 1e-3) / 1e9 if step_ms > 0 else 0.0,
                     "bpr_only_note": "SURVEY 8d's 24d+20 B/triple over the same duration (the figure to compare with non-fused kernels)",
                     "zero_grad_row_steps_replayed_per_step": row_steps_replayed / max(a.steps, 1),
                     "pre_aging_steps": age + PARITY_STEPS,
                     "variant": "bulk-copy pipeline (pda_step_pipe.cu)" if (d == 128 and os.environ.get("PDA_STEP_PIPE", "1") != "0")
                     else "register gather (pda_train.cu)"},
        "adam_apply": {"ms_per_step": adam_ms / max(a.steps, 1), "launches_per_step": adam_n / max(a.steps, 1),
                       "share_of_step": adam_ms / ms, "achieved": adam_gbs, "frac": adam_gbs / pk["hbm"],
                       "algorithmic_bytes": adam_bytes, "rows_per_step": adam_rows, "mode": a.adam, "unit": "GB/s"},
        "adam_catchup": {"ms_per_step": cat_ms / max(a.steps, 1), "share_of_step": cat_ms / ms},
        # the sampler of step k+1 runs on a side stream under step k's Adam sweep / exchange: its event pair spans the time
        # it spends queued behind those kernels' CTAs, so its elapsed time is not a share of the step (alone: 0.15 ms)
        "sampler": {"ms_per_launch": samp_ms / max(samp_n, 1), "concurrent_with_previous_step": True, "share_of_step": None},
    }
    dom = "adam_apply" if adam_ms > step_ms else "bpr_step"
    default_shape = (a.users, a.items, d, B, world, a.adam) == (10_000_000, 1_000_000, 128, 1 << 20, 1, "lazy_users")
    traffic = ncu_traffic() if default_shape else {}       # the capture was taken on exactly this workload
    for k in ("bpr_step", "adam_apply", "sampler"):
        kern[k]["traffic"] = traffic.get(k)
    roof = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved"], "peak": pk["hbm"], "unit": "GB/s",
            "frac": kern[dom]["frac"], "traffic": traffic.get(dom), "traffic_source": traffic.get("_file"),
            "algorithmic_bytes": kern[dom]["algorithmic_bytes"], "ms_per_launch": kern[dom].get("ms_per_launch"),
            "peak_source": pk["src"] + ", burst copy figure (kernel timed alone with CUDA events)"}

    # ---- CPU legs (rank 0, N = 1 only): the oracle's C port timed on the host cores + the parity checks ----
    cpu, parity = None, None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu, cpu_first = cpu_baseline(a, ds, P, users_local, steps=max(a.cpu_steps, PARITY_STEPS - 1))
        parity = train_parity(gpu_first, cpu_first)
        if ev is not None and ev_cpu_inputs is not None:
            ev["cpu_baseline"], ev_par = eval_cpu_leg(a, ev_cpu_inputs)
            parity.update(ev_par)
        parity["all_true"] = all(v for k, v in parity.items() if isinstance(v, bool))
    elif world > 1:
        parity = par_multi

    if rank == 0:
        clocks.stop()
        out = {"metric": "bpr_triples_per_sec", "value": value, "unit": "triples/s", "n_gpus": world, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": make_config(a, world), "adam_evaluation": a.adam + " (exact lazy replay of the every-row sweep: bit-identical tables)",
               "roofline": roof, "kernels": kern, "exchange": exch, "cpu_baseline": cpu, "e2e": e2e, "eval": ev, "parity": parity,
               "gpu_launches": int(step_n + adam_n + cat_n + samp_n + a.steps), "clocks": clk,
               "last_loss": [float(x) for x in loss]}
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------
def run_eval(a, model, ds, dev, world, rank, pk):
    """every rank scores eval_users of ITS user shard against all (replicated) items: no exchange on the data path
    (pda_b200.parallel.ShardedEvaluator adds the one metric-sum all-reduce); aggregate = world x pairs / max time."""
    import torch
    import torch.distributed as dist
    from pda_b200 import synth
    d = a.dim
    users_local = a.users // world
    Me = min(a.eval_users, users_local)
    eu = np.arange(Me, dtype=np.int32)
    tr = ncu_traffic()

    def one(pop_e, tag):
        model.do_recommendation(eu[:256], None, "condition", pos_pop=pop_e, K=50)     # warm-up
        model.profile(True)
        w0 = time.perf_counter()
        ids, sc = model.do_recommendation(eu, None, "condition", pos_pop=pop_e, K=50, return_scores=True)
        wall = time.perf_counter() - w0
        pr = model.profile_read()
        model.profile(False)
        kms = pr["eval_exact"][0] + pr["eval_tensor"][0]
        if world > 1:
            tk = torch.tensor([kms, wall], device=dev, dtype=torch.float64)
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
            kms_max, wall = float(tk[0].item()), float(tk[1].item())
        else:
            kms_max = kms
        pairs = Me * a.items                                  # per rank
        flop = pairs * 2 * d                                  # 2*d FLOP per scored (user, item) pair
        sb_ms, sb_n = pr["eval_sweep_b"]
        sa_ms, sa_n = pr["eval_sweep_a"]
        st = model.tc_last_stats() if pr["eval_tensor"][1] else None
        same = (a.items, d) == (1_000_000, 128)
        tf = flop / (kms * 1e-3) / 1e12
        r = {"tables": tag, "value": world * pairs / (kms_max * 1e-3), "unit": "pairs/s", "e2e_value": world * pairs / wall,
             "kernel_ms": kms,
             "filter_stats": st,
             "candidates_per_row": (st["candidates"] / max(st["rows"], 1)) if st else None,
             "rows_exact_fallback": st["rows_exact_fallback"] if st else None,
             "exact_fallback_ms": pr["eval_exact"][0],
             # whole pipeline (prep + sampled sweep + tau + full sweep + rescoring), rank 0.  The group of kernels is timed
             # once, in isolation -> the burst bf16 peak is the denominator; the sustained one is given beside it
             "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": tf / pk["bf16"],
                          "frac_of_sustained": tf / pk["bf16_sus"], "traffic": None,
                          "peak_source": pk["src"] + " burst bf16 (kernel group timed in isolation)"}}
        if sb_n:
            # the dominant eval kernel alone: the full sweep performs exactly the 2*d FLOP per pair once
            r["sweep_pass_b"] = {"ms": sb_ms, "launches": sb_n, "share_of_eval": sb_ms / kms,
                                 "achieved": flop / (sb_ms * 1e-3) / 1e12, "peak": pk["bf16"], "unit": "TFLOP/s",
                                 "frac": flop / (sb_ms * 1e-3) / 1e12 / pk["bf16"],
                                 "frac_of_sustained": flop / (sb_ms * 1e-3) / 1e12 / pk["bf16_sus"],
                                 "traffic": tr.get("eval_sweep_pass_b_per_%d_users" % min(Me, 32768)) if same else None,
                                 "mma": "tcgen05.mma kind::f16 128x128x16, bf16 -> fp32; K = d + 16 (extra block carries pop)"}
            r["sweep_pass_a"] = {"ms": sa_ms, "launches": sa_n, "share_of_eval": sa_ms / kms}
            r["other_kernels_ms"] = kms - sb_ms - sa_ms - pr["eval_exact"][0]
        return r, ids, sc

    out = {"metric": "eval_user_item_pairs_per_sec", "unit": "pairs/s", "n_gpus": world, "users_per_gpu": Me, "items": a.items,
           "K": 50, "rec_type": "condition",
           "backend": "tcgen05 bf16 filter (kind::f16, fp32 accumulate in TMEM, pop folded into the GEMM) + exact fp32 rescoring "
                      "of the certified candidates"}
    cpu_inputs = None
    if a.eval_tables in ("trained", "both"):
        pop_e = synth.eval_pop_torch(ds["pop"], GAMMA).cpu().numpy()
        r, ids, sc = one(pop_e, "trained: the tables after this run's Adam steps from Xavier init (|s| << 1: ranking ~ popularity, few candidates)")
        out["trained"] = r
    if a.eval_tables in ("fitted", "both"):
        # a fitted model's score spread: N(0,1)/sqrt(d) rows, skewed item norms, pop^6 (tests/test_gpu_eval.py::
        # test_recommend_tensor_large_item_set_sampled_pass) -- written into the first Me user rows and all item rows
        g = torch.Generator(device=dev); g.manual_seed(1234)
        Uv = dev_view(model.table_ptr("user_embedding"), (users_local, d), dev)
        Iv = dev_view(model.table_ptr("item_embedding"), (a.items, d), dev)
        Uv[:Me] = torch.randn((Me, d), device=dev, generator=g) / math.sqrt(d)
        Iv[:] = torch.randn((a.items, d), device=dev, generator=g) / math.sqrt(d)
        Iv *= (0.2 + torch.rand((a.items, 1), device=dev, generator=g) ** 4) * 3.0
        pop_f = (torch.rand(a.items, device=dev, generator=g) ** 6).float().cpu().numpy()
        torch.cuda.synchronize()
        r, ids, sc = one(pop_f, "fitted-like: N(0,1)/sqrt(d) rows, item norms x 3(0.2+u^4), pop = u^6")
        out["fitted"] = r
        pop_e = pop_f
    head = out.get("fitted") or out.get("trained")
    for k in ("value", "e2e_value", "kernel_ms", "roofline", "filter_stats"):
        out[k] = head[k]
    out["headline_tables"] = head["tables"]
    if rank == 0 and world == 1 and not a.no_cpu:
        # inputs of the CPU leg: the tables exactly as the GPU just scored them (the last regime run)
        n_cpu = int(min(2048, max(512, 64 * (os.cpu_count() or 8))))
        half = n_cpu // 2
        blocks = [np.arange(0, min(Me, 32768))]
        if Me > 32768:
            blocks.append(np.arange(32768, Me))
        sel = np.concatenate([b[np.linspace(0, len(b) - 1, half if len(blocks) == 2 else n_cpu).astype(np.int64)] for b in blocks])
        sel = np.unique(sel).astype(np.int32)
        Uv = dev_view(model.table_ptr("user_embedding"), (users_local, d), dev)
        U_sel = Uv[torch.as_tensor(sel.astype(np.int64), device=dev)].cpu().numpy()
        I_all = model.get_table("item_embedding")
        ip = ds["indptr"][torch.as_tensor(np.concatenate([sel, sel + 1]).astype(np.int64), device=dev)].cpu().numpy()
        lo, hi = ip[:len(sel)], ip[len(sel):]
        items_all = ds["items"]
        rows = [items_all[int(l):int(h)].cpu().numpy() for l, h in zip(lo, hi)]
        sub_ptr = np.zeros(len(sel) + 1, dtype=np.int64)
        sub_ptr[1:] = np.cumsum([len(r) for r in rows])
        cpu_inputs = dict(sel=sel, U=U_sel, I=I_all, indptr=sub_ptr, items=np.concatenate(rows).astype(np.int32), pop=pop_e,
                          gpu_ids=ids[sel], gpu_scores=sc[sel], tables=head["tables"], n_blocks=len(blocks))
    return out, cpu_inputs


def eval_cpu_leg(a, x):
    """C port of the all-items recommender (MF/train_new_api.py:594-612) on a bounded sample of the eval users, all host
    cores: timing = eval.cpu_baseline, result = the parity check of the tcgen05 path at the benchmarked shape."""
    from oracle import c_oracle as co
    co.build()
    cores = co.set_num_threads(os.cpu_count() or 1)
    n = len(x["sel"])
    IT = np.ascontiguousarray(x["I"].T)
    t0 = time.perf_counter()
    rid, rsc = co.recommend(x["U"], None, np.arange(n, dtype=np.int32), "condition", 50, x["indptr"], x["items"], pop=x["pop"], IT=IT)
    dt = time.perf_counter() - t0
    pairs = n * x["I"].shape[0]
    cpu = {"value": pairs / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
           "sample": f"{n} of the eval users (spread over {x['n_blocks']} blocks of 32768) x {x['I'].shape[0]} items, d={a.dim}, "
                     f"condition, mask + top-50; C/OpenMP restatement of MF/train_new_api.py:594-612", "seconds": dt}
    ids_eq = bool(np.array_equal(rid, x["gpu_ids"]))
    bits_eq = bool(np.array_equal(rsc.view(np.uint32), x["gpu_scores"].view(np.uint32)))
    par = {"eval_users_checked": n, "eval_user_blocks": x["n_blocks"], "eval_tables": x["tables"],
           "eval_topk_ids_equal_oracle": ids_eq, "eval_topk_score_bits_equal_oracle": bits_eq,
           "eval_rows_differing": int((rid != x["gpu_ids"]).any(axis=1).sum())}
    return cpu, par


def train_parity(gpu_first, cpu_first):
    """the first PARITY_STEPS steps of the GPU run against the C port on the same Philox batches (B = the bench's batch)"""
    g0, c0 = gpu_first["step0"], cpu_first["step0"]
    idx_eq = all(bool(np.array_equal(g0[k], c0[k])) for k in ("users", "pos", "neg"))
    pop_eq = all(bool(np.array_equal(g0[k].view(np.uint32), c0[k].view(np.uint32))) for k in ("pos_pop", "neg_pop"))
    gl = np.asarray(gpu_first["losses"], dtype=np.float64)
    cl = np.asarray(cpu_first["losses"][:len(gl)], dtype=np.float64)
    rel = float(np.max(np.abs(gl - cl) / np.maximum(np.abs(cl), 1e-30)))
    return {"step0_indices_bit_equal_oracle": idx_eq, "step0_pops_bit_equal_oracle": pop_eq, "train_steps_checked": int(len(gl)),
            "train_loss3_max_rel_err_vs_oracle": rel, "train_loss3_within_1e-5": bool(rel <= 1e-5),
            "gpu_losses": gl.tolist(), "oracle_losses": cl.tolist()}


def multi_gpu_parity(world, rank, local, dev, model, trainer):
    """N > 1 (untimed): (a) the item-table replicas are bit-identical on all ranks after the timed steps;
    (b) on a small problem, 3 sharded steps (this ShardedTrainer code path, device sampler) == ONE process stepping on
    the union batch (a second single-GPU model on rank 0 fed the gathered batches)."""
    import torch
    import torch.distributed as dist
    import pda_b200
    from pda_b200 import synth
    from pda_b200.parallel import ShardedTrainer
    trainer.finish()
    torch.cuda.synchronize()
    Iv = dev_view(model.table_ptr("item_embedding"), (model.n_items, model.emb_dim), dev, "<i4")
    chk = torch.stack([Iv.to(torch.int64).sum(), (Iv.to(torch.int64) * (torch.arange(Iv.shape[1], device=dev) + 1)).sum()])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    replicas_equal = all(bool(torch.equal(allc[0], c)) for c in allc)
    # (b)
    nu, ni, d, B, T = 20000, 4096 * world, 128, 2048, 10
    ds = synth.make_synthetic(nu, ni, seed=77 + rank, device=dev, mean_extra_deg=6.0, min_deg=4)
    P = synth.train_pop_matrix_torch(ds["pop"], GAMMA).cpu().numpy()
    m = pda_b200.PDAModel(nu, ni, d, train="s_condition", batch_size=B * world, lr=1e-2, regs=1e-3, device=local, max_batch=B,
                          seed=SEED_INIT, init=True)
    m.set_train_csr_device(wrap_ptr(ds["indptr"]), wrap_ptr(ds["items"]), wrap_ptr(ds["times"]), ds["nnz"],
                           wrap_ptr(ds["active"]), ds["active"].numel(), unique_times=np.arange(T - 1))
    # every rank's popularity table differs (its own synthetic shard): use rank 0's everywhere, like a replicated pop table
    Pl = [None] * world
    dist.all_gather_object(Pl, P if rank == 0 else None)
    P = Pl[0]
    m.set_train_pop(P)
    U0 = m.get_table("user_embedding")
    tr = ShardedTrainer(m, world, rank)
    stream = torch.cuda.current_stream().cuda_stream
    batches, losses = [], []
    for s in range(3):
        b = m.sample_batch(900 + rank, 0, s, B)
        batches.append({k: b[k] for k in ("users", "pos", "neg", "pos_pop", "neg_pop")})
        tr.train_sampled(900 + rank, 0, s, 1, B, stream)
        losses.append(m.read_loss(stream))
    tr.finish()
    Ur, Ir = m.get_table("user_embedding"), m.get_table("item_embedding")
    gathered = [None] * world
    dist.all_gather_object(gathered, (batches, Ur, U0))
    res = None
    if rank == 0:
        one = pda_b200.PDAModel(nu * world, ni, d, train="s_condition", batch_size=B * world, lr=1e-2, regs=1e-3, device=local,
                                max_batch=B * world, seed=SEED_INIT, init=True)
        one.set_table("user_embedding", np.concatenate([g[2] for g in gathered]))
        one.set_adam_mode("dense")
        ref_losses = []
        for s in range(3):
            cat = [np.concatenate([g[0][s][k] + (r * nu if k == "users" else 0) for r, g in enumerate(gathered)])
                   for k in ("users", "pos", "neg", "pos_pop", "neg_pop")]
            ref_losses.append(one.train_step(*cat))
        Uo, Io = one.get_table("user_embedding"), one.get_table("item_embedding")
        one.close()
        item_err = float(np.abs(Ir - Io).max() / np.abs(Io).max())
        user_err = max(float(np.abs(g[1] - Uo[r * nu:(r + 1) * nu]).max() / np.abs(Uo).max()) for r, g in enumerate(gathered))
        loss_err = float(np.max(np.abs(np.asarray(losses) - np.asarray(ref_losses)) / np.abs(np.asarray(ref_losses))))
        res = {"item_table_replicas_bit_identical": replicas_equal,
               "union_batch_check": {"shape": f"{world} x {nu} users, {ni} items, d={d}, B={B}/rank, 3 steps, device sampler",
                                     "item_table_max_err_of_scale": item_err, "user_table_max_err_of_scale": user_err,
                                     "loss3_max_rel_err": loss_err},
               "sharded_step_equals_union_batch": bool(item_err <= 1e-5 and user_err <= 1e-5 and loss_err <= 1e-5)}
        res["all_true"] = bool(res["item_table_replicas_bit_identical"] and res["sharded_step_equals_union_batch"])
    m.close()
    return res


# ------------------------------------------------------------------------------------------------------------
def cpu_baseline(a, ds, P, users_local, steps=8):
    """The oracle's C port (reference semantics, OpenMP over all host cores) on the same workload, from the same Philox
    init and batches as the GPU run: step 0 is the warm-up, steps 1..steps are timed; every loss is kept for the parity check."""
    from oracle import c_oracle as co
    co.build()
    cores = co.set_num_threads(os.cpu_count() or 1)
    B, d = a.batch, a.dim
    indptr = ds["indptr"].cpu().numpy(); items = ds["items"].cpu().numpy(); times = ds["times"].cpu().numpy()
    active = ds["active"].cpu().numpy()
    Pn = P.cpu().numpy()
    U = co.xavier_init(users_local, d, SEED_INIT, 0)
    I = co.xavier_init(a.items, d, SEED_INIT, 1)
    ref = co.CModel(U, I, a.lr, a.regs, B, "s_condition", copy=False)
    ut = np.arange(ds["n_stages"] - 1)
    first = {"losses": []}

    def one(s):
        b = co.sample_batch(SEED_SAMPLER, 0, s, B, active, indptr, items, times, a.items, ut, Pn)
        if s == 0:
            first["step0"] = {k: b[k].copy() for k in ("users", "pos", "neg", "pos_pop", "neg_pop")}
        first["losses"].append([float(x) for x in ref.train_step(b["users"], b["pos"], b["neg"], b["pos_pop"], b["neg_pop"])])
    one(0)
    t0 = time.perf_counter()
    for s in range(1, 1 + steps):
        one(s)
    dt = time.perf_counter() - t0
    return ({"value": steps * B / dt, "unit": "triples/s", "cores": cores, "kind": "port",
             "sample": f"{steps} full steps (B={B}, sampler + dense Adam over {users_local}+{a.items} rows) after 1 warm-up; "
                       "C/OpenMP restatement of the TF1 graph (TF1 itself cannot run here)",
             "ms_per_step": dt / steps * 1e3}, first)


def run_reference(a):
    """--impl reference: the CPU restatement timed on ALL host cores, always on the full workload (one rank's view of
    the config: users x items, B triples per step), whatever N is (TF1.14 is not installable here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from pda_b200 import synth
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ds = synth.make_synthetic(a.users, a.items, seed=SEED_DATA, device=dev)
    P = synth.train_pop_matrix_torch(ds["pop"], GAMMA)
    steps = max(1, min(a.steps, 25))          # 25 x ~0.4 s: the whole arm ends within a minute
    cpu, _ = cpu_baseline(a, ds, P, a.users, steps=steps)
    out = {"impl": "reference", "metric": "bpr_triples_per_sec", "value": cpu["value"], "unit": "triples/s",
           "n_gpus": a.gpus, "steps": steps, "warmup": 1, "ms_per_step": cpu["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": make_config(a, world), "adam_evaluation": "dense (the reference's own sweep)",
           "reference_arm": "all host cores on the full user table (users x items of the config, B triples per step), independent of N",
           "cpu_baseline": cpu,
           "e2e": {"value": cpu["value"], "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    args = parse_args()
    if args.config != "synthetic":
        from tools import bench_configs
        bench_configs.main(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
