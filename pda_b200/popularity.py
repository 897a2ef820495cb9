"""Popularity tables of PDA (host side, float64 numpy like the reference; one-off work outside the hot path).

Mirrors: pop_pre.py:12-57 (t_k.txt -> item_pop_seq_ori2.txt), MF/train_new_api.py:862-880 (load_popularity),
:882-893 (get_dataset_tot_popularity, PDG), :895-906 (get_popularity_from_load), :952-959 + :965-970 (eval
popularities).  float64 `pow` on the host and an fp32 cast at the device boundary is exactly what the reference
does (np.power in the driver, fp32 at tf.data's from_generator), so the tables are bit-identical inputs.
"""
from __future__ import annotations

import os

import numpy as np


def load_popularity(args):
    r_path = args.data_path + args.dataset + "/"
    pop_save_path = r_path + "item_pop_seq_ori2.txt"
    if not os.path.exists(pop_save_path):
        pop_save_path = r_path + "item_pop_seq_ori.txt"
    cache = r_path + "pda_cache_pop.npy"
    if not os.path.exists(pop_save_path) and not os.path.exists(cache):
        # the interaction loader falls back the other way round (data.py); be as forgiving here
        alt = "./data/" + args.dataset + "/"
        for cand in (alt + "item_pop_seq_ori2.txt", alt + "item_pop_seq_ori.txt"):
            if os.path.exists(cand):
                pop_save_path = cand
        cache = alt + "pda_cache_pop.npy" if os.path.exists(alt + "pda_cache_pop.npy") else cache
    print("popularity used:", pop_save_path)
    if os.path.exists(pop_save_path):
        print("pop save path: ", pop_save_path)
        raw = np.loadtxt(pop_save_path, dtype=np.float64, ndmin=2)
        pop_item_all = raw[:, 1:]          # rows are kept in FILE order, like the reference (item ids are not used)
    elif os.path.exists(cache):
        pop_item_all = np.load(cache)
    else:
        raise FileNotFoundError(pop_save_path)
    print("pop_item_all shape:", pop_item_all.shape)
    print("load pop information:", pop_item_all.mean(), pop_item_all.max(), pop_item_all.min())
    return pop_item_all


def pop_table_from_stage_files(path, n_stages, n_items):
    """pop_pre.py:12-57: per stage, (count+1)/(total+n_item) for listed items, 1/(total+n_item) otherwise,
    then min-max normalisation per stage.  Returns float64 [n_items, n_stages]."""
    out = np.zeros((n_items, n_stages), dtype=np.float64)
    for t in range(n_stages):
        cnt = np.zeros(n_items, dtype=np.float64)
        listed = np.zeros(n_items, dtype=bool)
        with open(os.path.join(path, "t_%d.txt" % t)) as f:
            for line in f:
                parts = line.split()
                if parts:
                    cnt[int(parts[0])] = len(parts) - 1
                    listed[int(parts[0])] = True
        total = cnt.sum()
        p = np.where(listed, (cnt + 1.0) / (total + n_items), 1.0 / (total + n_items))
        out[:, t] = (p - p.min()) / (p.max() - p.min())
    return out


def get_popularity_from_load(item_pop_all):
    popularity_matrix = item_pop_all[:, :-1]   # the last column is the test stage
    print("------ popularity information --------")
    print("   each stage mean:", popularity_matrix.mean(axis=0))
    print("   each stage max:", popularity_matrix.max(axis=0))
    print("   each stage min:", popularity_matrix.min(axis=0))
    return popularity_matrix


def get_dataset_tot_popularity(data):
    """global popularity of PDG (--train condition): (count+1) normalised to a distribution, then min-max."""
    cnt = np.bincount(data.train_items, minlength=data.n_items).astype(np.float64)
    p = cnt + 1.0
    p /= p.sum()
    p = (p - p.min()) / (p.max() - p.min())
    print("popularity information-- mean:{},max:{},min:{}".format(p.mean(), p.max(), p.min()))
    return p


def eval_popularities(pop_item_all, popularity_exp):
    """(last-stage pop) ** gamma and (linearly extrapolated pop, clipped to (1e-9, 1]) ** gamma
    -- prediction methods (a) and (b), train_new_api.py:952-959."""
    last = np.power(pop_item_all[:, -2], popularity_exp)
    lin = pop_item_all[:, -2] + 0.5 * (pop_item_all[:, -2] - pop_item_all[:, -3])
    lin[np.where(lin <= 0)] = 1e-9
    lin[np.where(lin > 1.0)] = 1.0
    lin = np.power(lin, popularity_exp)
    return last, lin


def bprmf_a_popularities(pop_item_all, linear_predict_popularity_powered):
    """--train normal (BPRMF-A): un-powered last-stage and linear popularity; the reference clips the linear
    one with masks computed from the *powered* array (train_new_api.py:967-970) -- kept."""
    last_ori = pop_item_all[:, -2]
    lin_ori = pop_item_all[:, -2] + 0.5 * (pop_item_all[:, -2] - pop_item_all[:, -3])
    lin_ori[np.where(linear_predict_popularity_powered <= 0)] = 1e-9
    lin_ori[np.where(linear_predict_popularity_powered > 1.0)] = 1.0
    return last_ori, lin_ori
