"""Times the fused NVLink-multicast exchange kernel (pda_dp_exchange_adam) alone, over grid sizes / unroll factors / debug
modes (multicast reads or writes switched off), on N GPUs:  torchrun --nproc-per-node N tools/dpx_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import pda_b200
from pda_b200.parallel import ShardedTrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_items, d = 1_000_000, 128
m = pda_b200.PDAModel(4096, n_items, d, train="s_condition", batch_size=1024, device=local, seed=2021)
P2P = os.environ.get("DPX_P2P", "0") == "1"
tr = ShardedTrainer(m, world, rank, exchange="p2p" if P2P else "nvls")
h = tr._nvls["hdl"]
lo, hi = tr._own[0]
variants = []
for blocks in ((148, 296, 592) if world > 2 else (32, 74, 148, 296, 592, 1184)):
    for unroll in ((2, 4) if world > 2 else (1, 2, 4, 8)):
        variants.append(dict(PDA_DPX_BLOCKS=blocks, PDA_DPX_UNROLL=unroll, PDA_DPX_DBG=3))
for dbg in (0, 1, 2):
    variants.append(dict(PDA_DPX_BLOCKS=592, PDA_DPX_UNROLL=4, PDA_DPX_DBG=dbg))
for dbg in (3 | 4, 3 | 8, 3 | 4 | 8):      # weak multimem stores / loads
    for unroll in (2, 8):
        variants.append(dict(PDA_DPX_BLOCKS=296, PDA_DPX_UNROLL=unroll, PDA_DPX_DBG=dbg))
variants.append(dict(PDA_DPX_BLOCKS=296, PDA_DPX_UNROLL=8, PDA_DPX_DBG=3))
variants.append(dict(PDA_DPX_BLOCKS=74, PDA_DPX_UNROLL=8, PDA_DPX_DBG=3))
variants.append(dict(PDA_DPX_BLOCKS=1184, PDA_DPX_UNROLL=1, PDA_DPX_DBG=3))
if P2P:
    variants = [dict(PDA_DPX_BLOCKS=b, PDA_DPX_UNROLL=u, PDA_DPX_P2P_MODE=md) for md in (0, 1, 2, 3) for b in (148, 592, 1184) for u in (2, 4, 8)]
for v in variants:
    os.environ.update({k: str(x) for k, x in v.items()})
    ts = []
    for it in range(6):
        h.barrier(channel=0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if P2P:
            m.dp_exchange_adam_p2p(tr._nvls["peerG"][0], tr._nvls["peerW"], rank, lo, hi, 0)
        else:
            m.dp_exchange_adam(tr._nvls["mcG"][0], tr._nvls["mcW"], lo, hi, 0)
        e1.record()
        h.barrier(channel=1)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([min(ts[1:])], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sl = (hi - lo) * d * 4 / 1e9
        print(json.dumps({**v, "ms": round(float(t.item()), 4), "slice_GB": sl, "in+out_GBps_per_dir": round(sl * world / (float(t.item()) * 1e-3), 1)}), flush=True)
m.close()
dist.destroy_process_group()
