"""Synthetic interaction sets in the reference's data shapes.

torch is plumbing here: it only manufactures the inputs (a user->items CSR with stage labels and the
per-stage popularity table of pop_pre.py:12-57) directly in HBM so a 10M x 1M set needs no host pass.
Nothing on the measured path runs through torch.
"""
from __future__ import annotations

import numpy as np


def make_synthetic(n_users, n_items, n_stages=10, mean_extra_deg=16.0, min_deg=8, seed=2020, device="cuda",
                   chunk_users=2_000_000):
    """Returns dict(indptr int64[U+1], items int32[nnz] (sorted within row), times uint8[nnz], active int32[],
    pop float64 [n_items, n_stages] (pop_pre.py formula; the last stage is the eval stage)) as torch tensors
    on `device`.  degree = min_deg + Poisson(mean_extra_deg); item ~ log-uniform rank over a fixed
    permutation (Zipf-like, exponent 1); stage uniform over the n_stages-1 train stages."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    deg = min_deg + torch.poisson(torch.full((n_users,), float(mean_extra_deg), device=device), generator=g).to(torch.int64)
    deg = torch.clamp(deg, max=max(1, n_items // 2))
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=indptr[1:])
    nnz = int(indptr[-1].item())
    perm = torch.randperm(n_items, device=device, generator=g).to(torch.int32)
    items = torch.empty(nnz, dtype=torch.int32, device=device)
    times = torch.empty(nnz, dtype=torch.uint8, device=device)
    log_n = float(np.log(n_items + 1.0))
    n_train = n_stages - 1
    for u0 in range(0, n_users, chunk_users):
        u1 = min(n_users, u0 + chunk_users)
        lo, hi = int(indptr[u0].item()), int(indptr[u1].item())
        n = hi - lo
        if n == 0:
            continue
        uid = torch.repeat_interleave(torch.arange(u0, u1, device=device), deg[u0:u1])
        r = torch.rand(n, device=device, generator=g, dtype=torch.float64)
        rank = torch.clamp(torch.exp(r * log_n).to(torch.int64) - 1, 0, n_items - 1)
        it = perm[rank].to(torch.int64)
        t = torch.randint(0, n_train, (n,), device=device, generator=g, dtype=torch.int64)
        key = (uid * n_items + it) * 16 + t            # sort by (user, item); stage rides along in the low bits
        key, _ = torch.sort(key)
        times[lo:hi] = (key & 15).to(torch.uint8)
        items[lo:hi] = ((key >> 4) % n_items).to(torch.int32)
        del uid, r, rank, it, t, key
    active = torch.nonzero(deg > 0).flatten().to(torch.int32)
    # pop_pre.py:31-42: (cnt+1)/(total+n_item), items absent from the stage -> 1/(total+n_item); min-max per stage
    cnt = torch.bincount(times.to(torch.int64) * n_items + items.to(torch.int64), minlength=n_train * n_items)
    cnt = cnt.view(n_train, n_items).to(torch.float64)
    cnt = torch.cat([cnt, cnt[-1:].clone()], 0)        # eval stage: reuse the last train stage's counts
    total = cnt.sum(1, keepdim=True)
    p = torch.where(cnt > 0, (cnt + 1.0) / (total + n_items), 1.0 / (total + n_items))
    pmin, pmax = p.min(1, keepdim=True).values, p.max(1, keepdim=True).values
    pop = ((p - pmin) / (pmax - pmin)).t().contiguous()
    return dict(indptr=indptr, items=items, times=times, active=active, pop=pop, nnz=nnz, n_stages=n_stages)


def train_pop_matrix_torch(pop, gamma):
    """(pop[:, :-1]) ** gamma in float64, cast to fp32 (train_new_api.py:988-990 + the fp32 feed at :552)."""
    import torch
    return torch.pow(pop[:, :-1], gamma).to(torch.float32).contiguous()


def eval_pop_torch(pop, gamma):
    """last-stage popularity ** gamma (train_new_api.py:952-953)."""
    import torch
    return torch.pow(pop[:, -2], gamma).to(torch.float32).contiguous()
