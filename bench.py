#!/usr/bin/env python
"""bench.py -- headline benchmark of the PDA hot path on B200 (contract: see DESIGN.md section 7).

  python bench.py [--gpus N] [--steps K] [--warmup W]                # this repo's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # CPU reference arm (oracle port)

A "step" = one pass of the hot path over one batch: device sampler -> fused BPR step kernel
(gather, dots, ELU'/pop^gamma, log-sigmoid BPR loss + L2, gradient scatter-add) -> TF1-semantics Adam
sweep of both tables.  Workload (BASELINE.json configs[4], fits one B200): synthetic 10M users x 1M
items, d=128, PD (--train s_condition), gamma=0.16.  `value` = triples/s with everything resident in
HBM; `e2e` = the same metric through PDAModel.train_step (the sess.run-shaped host API: pinned host
batch -> H2D -> step -> D2H of the 3 loss scalars).
"""
from __future__ import annotations

import argparse
import json
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--users", type=int, default=10_000_000)
    ap.add_argument("--items", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=1 << 20, help="triples per step per GPU")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--regs", type=float, default=1e-3)
    ap.add_argument("--eval-users", type=int, default=65536, help="users scored against all items (blocks of 32768)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=10, help="steps of the cpu_baseline leg (about 0.4 s each on 16 cores)")
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
    """nvidia-smi clock / throttle-reason samples DURING the timed region (profiling recipe's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def wrap_ptr(t):
    return int(t.data_ptr())


def ncu_traffic():
    """DRAM bytes per launch of the kernels of THIS workload from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        return json.load(open(p))
    except (OSError, ValueError):
        return {}


def make_config(a, world):
    """identical for our arm and the reference arm (the driver compares the two lines)"""
    users_local = a.users // world
    return {"workload": f"synthetic {a.users} users x {a.items} items d={a.dim}, PD (s_condition) gamma={GAMMA}, "
                        f"TF1 every-row Adam semantics, B={a.batch} triples/step/GPU",
            "users": a.users, "items": a.items, "d": a.dim, "batch_per_gpu": a.batch, "global_batch": a.batch * world,
            "parallelism": f"user-shard x{world}, items replicated" + (" + NCCL item-grad allreduce" if world > 1 else ""),
            "l2_policy": "tables >> L2 (user table %.1f GB per rank): no flush needed" % (users_local * a.dim * 4 / 1e9)}


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
        # high-priority NCCL stream: the item-gradient all-reduce must get its CTAs although the (persistent, full-
        # occupancy) Adam kernel of the rank-local half becomes runnable at the same instant
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    dev = torch.device("cuda", local)
    pk = peaks()
    B, d = a.batch, a.dim

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
    trainer = ShardedTrainer(model, world, rank)      # world > 1: lazy user table + dense (all-reduced) item table
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run_steps(step0, n):
        trainer.train_sampled(SEED_SAMPLER + rank, 0, step0, n, B, stream)

    # ---- device-resident run: `value` ----
    run_steps(0, a.warmup)
    barrier()
    model.profile(True)
    if a.adam != "dense":
        model.adam_stats(reset=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_steps(a.warmup, a.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if rank == 0:
        clocks.stop()
    prof = model.profile_read()
    model.profile(False)
    rows_updated, row_steps_replayed = model.adam_stats(reset=True) if a.adam != "dense" else (0, 0)
    loss = model.read_loss(stream)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = a.steps * B * world / (ms * 1e-3)

    # ---- end-to-end through the host API: `e2e` ----
    e2e = None
    if not a.no_e2e:
        n_e2e = min(a.steps, 10)
        keys = ("users", "pos", "neg", "pos_pop", "neg_pop")
        # host batches live in pinned memory (cudaHostAlloc through the C ABI), [n, B] per array
        first = model.sample_batch(SEED_SAMPLER + rank, 1, 0, B)
        pin = {k: model.pinned_array((n_e2e + 2, B), first[k].dtype) for k in keys}
        for s in range(n_e2e + 2):
            b = first if s == 0 else model.sample_batch(SEED_SAMPLER + rank, 1, s, B)
            for k in keys:
                pin[k][s] = b[k]
        if world == 1:
            # the generator-fed epoch loop in one call: batch k+1's copies overlap step k (pda_train_steps_host)
            model.train_steps(*(pin[k][:2] for k in keys))
            barrier()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            w0 = time.perf_counter()
            t0.record()
            l3 = model.train_steps(*(pin[k][2:] for k in keys))[-1]
            t1.record()
            barrier()
            api = "PDAModel.train_steps (pda_train_steps_host: n pinned host batches, copies pipelined with the steps)"
        else:
            for s in range(2):
                trainer.train_step_host(*(pin[k][s] for k in keys))
            barrier()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            w0 = time.perf_counter()
            t0.record()
            for s in range(2, n_e2e + 2):
                l3 = trainer.train_step_host(*(pin[k][s] for k in keys))
            t1.record()
            barrier()
            api = "ShardedTrainer.train_step_host (pda_stage_batch_host + split step + NCCL exchange)"
        wall = time.perf_counter() - w0
        ems = max(t0.elapsed_time(t1), 0.0)
        te = torch.tensor([max(ems * 1e-3, wall)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": n_e2e * B * world / float(te.item()), "unit": "triples/s", "h2d_bytes_per_step": 20 * B,
               "d2h_bytes_per_step": 12, "steps": n_e2e, "api": api, "last_loss": [float(x) for x in l3]}

    # ---- eval: all-items scoring + pop adjust + mask + top-50 (pairs/s) ----
    ev = None
    if not a.no_eval:
        # every rank scores eval_users of ITS user shard against all (replicated) items: no exchange on the data path
        # (pda_b200.parallel.ShardedEvaluator adds the one metric-sum all-reduce); aggregate = world x pairs / max time
        Me = min(a.eval_users, users_local)
        eu = np.arange(Me, dtype=np.int32)
        pop_e = synth.eval_pop_torch(ds["pop"], GAMMA).cpu().numpy()
        model.do_recommendation(eu[:256], None, "condition", pos_pop=pop_e, K=50)     # warm-up
        model.profile(True)
        w0 = time.perf_counter()
        ids = model.do_recommendation(eu, None, "condition", pos_pop=pop_e, K=50)
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
        tr = ncu_traffic()
        same = (a.items, d) == (1_000_000, 128)
        ev = {"metric": "eval_user_item_pairs_per_sec", "value": world * pairs / (kms_max * 1e-3), "unit": "pairs/s",
              "e2e_value": world * pairs / wall, "n_gpus": world, "users_per_gpu": Me, "items": a.items, "K": 50,
              "rec_type": "condition",
              "backend": "exact fp32 CUDA-core scorer" if pr["eval_tensor"][1] == 0 else
              "tcgen05 bf16 filter (kind::f16, fp32 accumulate in TMEM, pop folded into the GEMM) + exact fp32 rescoring "
              "of the certified candidates",
              "filter_stats": model.tc_last_stats() if pr["eval_tensor"][1] else None,
              "kernel_ms": kms,
              # whole pipeline (prep + sampled sweep + tau + full sweep + rescoring) against the tensor peak, rank 0
              "roofline": {"bound": "tensor", "achieved": flop / (kms * 1e-3) / 1e12, "peak": pk["bf16_sus"],
                           "unit": "TFLOP/s", "frac": flop / (kms * 1e-3) / 1e12 / pk["bf16_sus"],
                           "traffic": None, "peak_source": pk["src"] + " sustained bf16 (kernels timed inside a long step)"}}
        if sb_n:
            # the dominant eval kernel alone: the full sweep performs exactly the 2*d FLOP per pair once
            ev["sweep_pass_b"] = {"ms": sb_ms, "launches": sb_n, "share_of_eval": sb_ms / kms,
                                  "achieved": flop / (sb_ms * 1e-3) / 1e12, "peak": pk["bf16"], "unit": "TFLOP/s",
                                  "frac": flop / (sb_ms * 1e-3) / 1e12 / pk["bf16"],
                                  "frac_of_sustained": flop / (sb_ms * 1e-3) / 1e12 / pk["bf16_sus"],
                                  "traffic": tr.get("eval_sweep_pass_b_per_%d_users" % min(Me, 32768)) if same else None,
                                  "peak_source": pk["src"] + " burst bf16 (kernel timed alone with CUDA events)",
                                  "mma": "tcgen05.mma kind::f16 128x128x16, bf16 -> fp32; K = d + 16 (extra block carries pop)"}
            ev["sweep_pass_a"] = {"ms": sa_ms, "launches": sa_n, "share_of_eval": sa_ms / kms}

    # ---- per-kernel device times (CUDA events on the launching stream) and rooflines ----
    step_ms, step_n = prof["bpr_step"]
    adam_ms, adam_n = prof["adam"]
    cat_ms, cat_n = prof["adam_catchup"]
    samp_ms, samp_n = prof["sampler"]
    bytes_triple = 24 * d + 20
    # lazy mode + distinct users (device sampler): the step kernel also owns the Adam update of its B user rows
    # (catch-up replay + apply) -> its algorithmic bytes are SURVEY 8d's step figure PLUS SURVEY 8d's
    # Adam-apply figure (rows x d x 4 x 6) for those B rows; the separate apply kernel then covers item rows only
    fused = a.adam in ("lazy", "lazy_users") and os.environ.get("PDA_FUSE_USER_ADAM", "1") != "0"
    step_bytes = bytes_triple * B + (B * d * 4 * 6 if fused else 0)
    step_gbs = step_bytes / (step_ms / max(step_n, 1) * 1e-3) / 1e9 if step_ms > 0 else 0.0
    # rows the Adam kernels outside the step kernel handle per step: dense-swept tables + lazily updated rows
    adam_rows = rows_updated / max(a.steps, 1)
    if a.adam == "dense":
        adam_rows += users_local + a.items
    elif a.adam == "lazy_users" or world > 1:
        adam_rows += a.items
    adam_bytes = adam_rows * d * 4 * 6                                       # W, m, v read + written
    adam_gbs = adam_bytes / (adam_ms / max(a.steps, 1) * 1e-3) / 1e9 if adam_ms > 0 else 0.0
    kern = {
        "bpr_step": {"ms_per_launch": step_ms / max(step_n, 1), "share_of_step": step_ms / ms, "achieved": step_gbs,
                     "frac": step_gbs / pk["hbm"], "algorithmic_bytes": step_bytes, "unit": "GB/s",
                     "bytes_model": "B*(24d+20)" + (" + B*d*4*6 (fused Adam of the B distinct user rows)" if fused else ""),
                     "bpr_only_gbs": bytes_triple * B / (step_ms / max(step_n, 1) * 1e-3) / 1e9 if step_ms > 0 else 0.0},
        "adam_apply": {"ms_per_step": adam_ms / max(a.steps, 1), "launches_per_step": adam_n / max(a.steps, 1),
                       "share_of_step": adam_ms / ms, "achieved": adam_gbs, "frac": adam_gbs / pk["hbm"],
                       "algorithmic_bytes": adam_bytes, "rows_per_step": adam_rows, "mode": a.adam, "unit": "GB/s"},
        "adam_catchup": {"ms_per_step": cat_ms / max(a.steps, 1), "share_of_step": cat_ms / ms,
                         "zero_grad_row_steps_replayed_per_step": row_steps_replayed / max(a.steps, 1)},
        # the sampler of step k+1 runs on a side stream under step k's Adam sweep: its event pair spans the time it spends
        # queued behind that kernel's CTAs, so its elapsed time is not a share of the step (alone it takes 0.15 ms)
        "sampler": {"ms_per_launch": samp_ms / max(samp_n, 1), "concurrent_with_previous_step": world == 1,
                    "share_of_step": None if world == 1 else samp_ms / ms},
    }
    dom = "adam_apply" if adam_ms > step_ms else "bpr_step"
    default_shape = (a.users, a.items, d, B, world, a.adam) == (10_000_000, 1_000_000, 128, 1 << 20, 1, "lazy_users")
    traffic = ncu_traffic() if default_shape else {}       # the capture was taken on exactly this workload
    for k in ("bpr_step", "adam_apply", "sampler"):
        kern[k]["traffic"] = traffic.get(k)
    roof = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved"], "peak": pk["hbm"], "unit": "GB/s",
            "frac": kern[dom]["frac"], "traffic": traffic.get(dom), "algorithmic_bytes": kern[dom]["algorithmic_bytes"],
            "peak_source": pk["src"] + ", burst copy figure (kernel timed alone with CUDA events)"}

    cpu = None
    if rank == 0 and not a.no_cpu:
        cpu = cpu_baseline(a, ds, P, users_local, steps=a.cpu_steps)

    if rank == 0:
        out = {"metric": "bpr_triples_per_sec", "value": value, "unit": "triples/s", "n_gpus": world, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": dict(make_config(a, world), adam_evaluation=a.adam),
               "roofline": roof, "kernels": kern, "cpu_baseline": cpu, "e2e": e2e, "eval": ev,
               "gpu_launches": int(step_n + adam_n + cat_n + samp_n + a.steps), "clocks": clocks.summary(),
               "last_loss": [float(x) for x in loss]}
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------
def cpu_baseline(a, ds, P, users_local, steps=3):
    """The oracle's C port (reference semantics, OpenMP over all host cores) on the same workload."""
    from oracle import c_oracle as co
    co.build()
    B, d = a.batch, a.dim
    indptr = ds["indptr"].cpu().numpy(); items = ds["items"].cpu().numpy(); times = ds["times"].cpu().numpy()
    active = ds["active"].cpu().numpy()
    Pn = P.cpu().numpy()
    U = co.xavier_init(users_local, d, SEED_INIT, 0)
    I = co.xavier_init(a.items, d, SEED_INIT, 1)
    ref = co.CModel(U, I, a.lr, a.regs, B, "s_condition", copy=False)
    ut = np.arange(ds["n_stages"] - 1)
    def one(s):
        b = co.sample_batch(SEED_SAMPLER, 0, s, B, active, indptr, items, times, a.items, ut, Pn)
        return ref.train_step(b["users"], b["pos"], b["neg"], b["pos_pop"], b["neg_pop"])
    one(0)
    t0 = time.perf_counter()
    for s in range(1, 1 + steps):
        one(s)
    dt = time.perf_counter() - t0
    return {"value": steps * B / dt, "unit": "triples/s", "cores": co.num_threads(), "kind": "port",
            "sample": f"{steps} full steps (B={B}, dense Adam over {users_local}+{a.items} rows) after 1 warm-up; "
                      "C/OpenMP restatement of the TF1 graph (TF1 itself cannot run here)",
            "ms_per_step": dt / steps * 1e3}


def run_reference(a):
    """--impl reference: the CPU restatement timed on the host cores (TF1.14 is not installable here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from pda_b200 import synth
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    users_local = a.users // world
    ds = synth.make_synthetic(users_local, a.items, seed=SEED_DATA, device=dev)
    P = synth.train_pop_matrix_torch(ds["pop"], GAMMA)
    steps = max(1, min(a.steps, 25))          # 25 x ~0.4 s: the whole arm ends within a minute
    cpu = cpu_baseline(a, ds, P, users_local, steps=steps)
    out = {"impl": "reference", "metric": "bpr_triples_per_sec", "value": cpu["value"], "unit": "triples/s",
           "n_gpus": a.gpus, "steps": steps, "warmup": 1, "ms_per_step": cpu["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": dict(make_config(a, world), adam_evaluation="dense (the reference's own sweep)"),
           "cpu_baseline": cpu,
           "e2e": {"value": cpu["value"], "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
