"""Times variants of the fused BPR step kernel on the bench workload inside ONE process (the kernel knobs are read per
launch): register-gather kernel vs the bulk-copy pipeline at several ring depths / CTA shapes / L2 policies.
usage: python tools/step_variants.py [--users 10000000] [--items 1000000] [--steps 40]   (GPU only; prints one JSON per variant)"""
import argparse
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=10_000_000)
    ap.add_argument("--items", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--variants", default="")
    a = ap.parse_args()
    import torch
    import pda_b200
    from pda_b200 import synth
    dev = torch.device("cuda", 0)
    d, B = 128, a.batch
    ds = synth.make_synthetic(a.users, a.items, seed=2020, device=dev)
    P = synth.train_pop_matrix_torch(ds["pop"], 0.16)
    m = pda_b200.PDAModel(a.users, a.items, d, train="s_condition", batch_size=B, lr=1e-3, regs=1e-3, max_batch=B, seed=2021)
    m.set_train_csr_device(ds["indptr"].data_ptr(), ds["items"].data_ptr(), ds["times"].data_ptr(), ds["nnz"],
                           ds["active"].data_ptr(), ds["active"].numel(), unique_times=np.arange(ds["n_stages"] - 1))
    m.set_train_pop(P.cpu().numpy())
    step = 0
    age = int(math.ceil(3.0 * a.users / B))
    m.train_sampled(2020, 0, step, age, B); step += age
    m.synchronize()
    variants = [dict(PDA_STEP_PIPE="0"),
                dict(PDA_STEP_PIPE="1", PDA_STEP_PIPE_D="3", PDA_STEP_PIPE_NW="8", PDA_STEP_PIPE_HINTS="0"),
                dict(PDA_STEP_PIPE="1", PDA_STEP_PIPE_D="3", PDA_STEP_PIPE_NW="8", PDA_STEP_PIPE_HINTS="1"),
                dict(PDA_STEP_PIPE="1", PDA_STEP_PIPE_D="3", PDA_STEP_PIPE_NW="8", PDA_STEP_PIPE_HINTS="2"),
                dict(PDA_STEP_PIPE="1", PDA_STEP_PIPE_D="3", PDA_STEP_PIPE_NW="8", PDA_STEP_PIPE_HINTS="3"),
                dict(PDA_STEP_PIPE="1", PDA_STEP_PIPE_D="2", PDA_STEP_PIPE_NW="8", PDA_STEP_PIPE_HINTS="1"),
                dict(PDA_STEP_PIPE="1", PDA_STEP_PIPE_D="4", PDA_STEP_PIPE_NW="8", PDA_STEP_PIPE_HINTS="1"),
                dict(PDA_STEP_PIPE="1", PDA_STEP_PIPE_D="6", PDA_STEP_PIPE_NW="8", PDA_STEP_PIPE_HINTS="1"),
                dict(PDA_STEP_PIPE="1", PDA_STEP_PIPE_D="3", PDA_STEP_PIPE_NW="4", PDA_STEP_PIPE_HINTS="1"),
                dict(PDA_STEP_PIPE="1", PDA_STEP_PIPE_D="4", PDA_STEP_PIPE_NW="4", PDA_STEP_PIPE_HINTS="1")]
    if a.variants:
        variants = [dict(kv.split("=") for kv in v.split(",")) for v in a.variants.split(";")]
    top_items = torch.argsort(torch.bincount(ds["items"].to(torch.int64), minlength=a.items), descending=True)[:28].cpu().numpy()
    for v in variants:
        v = dict(v)
        if "HOT" in v:      # popular-item rows pre-summed in shared memory: how many (0 = none; default = the library's own list)
            m.set_hot_items(top_items[:int(v["HOT"])])
        for k in ("PDA_STEP_PIPE", "PDA_STEP_PIPE_D", "PDA_STEP_PIPE_NW", "PDA_STEP_PIPE_HINTS", "PDA_STEP_PIPE_V"):
            os.environ.pop(k, None)
        os.environ.update({k: x for k, x in v.items() if k != "HOT"})
        m.train_sampled(2020, 0, step, 5, B); step += 5
        m.synchronize()
        m.profile(True)
        m.adam_stats(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.train_sampled(2020, 0, step, a.steps, B); step += a.steps
        e1.record()
        torch.cuda.synchronize()
        pr = m.profile_read()
        m.profile(False)
        _, replayed = m.adam_stats(reset=True)
        sms = pr["bpr_step"][0] / max(pr["bpr_step"][1], 1)
        print(json.dumps({"variant": v, "ms_per_step": e0.elapsed_time(e1) / a.steps, "bpr_step_ms": sms,
                          "bpr_step_gbs_40d20": B * (40 * d + 20) / (sms * 1e-3) / 1e9,
                          "adam_ms": pr["adam"][0] / max(pr["adam"][1], 1), "replayed_per_step": replayed / a.steps,
                          "loss": m.read_loss()}), flush=True)
    m.close()


if __name__ == "__main__":
    main()
