"""Shared builders for the test-suite (synthetic interaction sets in the reference's shapes)."""
import numpy as np


def synth_interactions(n_users, n_items, mean_deg, n_stages, seed, empty_frac=0.0):
    """Random (user, item, stage) interactions, unique (user, item) pairs, Zipf-ish item draw."""
    rng = np.random.default_rng(seed)
    deg = 1 + rng.poisson(mean_deg, n_users)
    if empty_frac > 0:
        deg[rng.random(n_users) < empty_frac] = 0
    deg = np.minimum(deg, n_items // 2)
    uid = np.repeat(np.arange(n_users), deg)
    w = 1.0 / (1.0 + np.arange(n_items))
    w /= w.sum()
    iid = rng.choice(n_items, size=len(uid), p=w)
    key = uid.astype(np.int64) * n_items + iid
    _, first = np.unique(key, return_index=True)
    first.sort()
    uid, iid = uid[first], iid[first]
    t = rng.integers(0, n_stages, len(uid))
    return uid, iid, t


def pop_table(n_items, n_stages, seed):
    rng = np.random.default_rng(seed)
    p = rng.random((n_items, n_stages + 1)) ** 3
    p[rng.random(p.shape) < 0.2] = 0.0
    p /= p.max(axis=0, keepdims=True)
    return p
