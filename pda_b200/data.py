"""Interaction data of the reference's formats as CSR arrays (host side; numpy only).

Mirrors MF/load_data.py of the reference:
  Data   (:24-110)   train.txt / valid.txt / test.txt          = `uid iid iid ...` per line       (--train normal)
  Data2  (:617-710)  train_with_time.txt = `uid iid time stars` per line + valid.txt / test.txt  (PD / PDA / BPR(t)-pop)
Same attribute names (n_users, n_items, n_train, n_valid, n_test, train_user_list, train_user_list_time,
valid_user_list, test_user_list, unique_times, users, items, batch_size, add_expo_popularity) so the driver
reads like the reference's; the storage behind them is CSR (int64 indptr, int32 items sorted within the row,
uint8 stage) because that is what the device sampler and the eval mask consume.

Path rule (reference quirk): interaction files are read from ./data/<dataset>/ whatever
--data_path says (load_data.py:27,619); if that directory does not exist, <data_path>/<dataset>/ is tried.
A parsed copy is cached next to the text files (pda_cache_<kind>.npz); a directory that holds only the cache
(the GPU boxes get exactly that, see tools/stage_douban.py) loads from it.
"""
from __future__ import annotations

import os

import numpy as np

CACHE_VERSION = 2


class CsrDict:
    """Read-only dict-of-lists view over CSR rows (only rows with at least one entry are keys, in `order`)."""

    def __init__(self, indptr, values, order=None):
        self.indptr, self.values = indptr, values
        deg = np.diff(indptr)
        self._keys = np.nonzero(deg > 0)[0] if order is None else np.asarray(order, dtype=np.int64)

    def __getitem__(self, u):
        u = int(u)
        if u < 0 or u + 1 >= len(self.indptr):
            return []
        return self.values[self.indptr[u]:self.indptr[u + 1]].tolist()

    def get(self, u, default=None):
        r = self[u]
        return r if r else default

    def __contains__(self, u):
        u = int(u)
        return 0 <= u < len(self.indptr) - 1 and self.indptr[u + 1] > self.indptr[u]

    def keys(self):
        return self._keys.tolist()

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._keys)

    def items(self):
        return ((int(u), self[u]) for u in self._keys)


def _parse_user_lists(path):
    """`uid iid iid ...` lines -> (uids in file order, counts, flat item array).  Lines without items are
    skipped like load_data.py:60-61; a later line of the same user replaces the earlier one (dict assignment)."""
    with open(path) as f:
        txt = f.read()
    lines = [ln for ln in txt.split("\n") if ln.strip()]
    ntok = np.fromiter((len(ln.split()) for ln in lines), dtype=np.int64, count=len(lines))
    flat = np.array(txt.split(), dtype=np.int64)
    assert flat.size == ntok.sum()
    starts = np.concatenate([[0], np.cumsum(ntok)[:-1]])
    keep = ntok > 1
    uids = flat[starts[keep]]
    counts = ntok[keep] - 1
    sel = np.ones(flat.size, dtype=bool)
    sel[starts] = False
    if (~keep).any():   # a bare `uid` line: its only token is already dropped by `sel`
        pass
    items = flat[sel]
    # duplicates of a uid: the last line wins
    if len(np.unique(uids)) != len(uids):
        last = {}
        off = np.concatenate([[0], np.cumsum(counts)])
        for k, u in enumerate(uids.tolist()):
            last[u] = k
        order = sorted(last.values())
        items = np.concatenate([items[off[k]:off[k + 1]] for k in order]) if order else items[:0]
        uids, counts = uids[order], counts[order]
    return uids, counts, items


def _csr_from_lists(n_rows, uids, counts, items, sort_rows):
    deg = np.zeros(n_rows, dtype=np.int64)
    deg[uids] = counts
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(deg, out=indptr[1:])
    row_of = np.repeat(uids, counts)
    if sort_rows:
        order = np.lexsort((np.arange(len(items)), items, row_of))
    else:
        order = np.argsort(row_of, kind="stable")
    return indptr, items[order].astype(np.int32), order


class _Interactions:
    kind = "?"

    def __init__(self, args):
        self.batch_size = args.batch_size
        self.dataset = args.dataset
        if getattr(args, "model", "mf") not in ("mf", "biasmf"):
            raise NotImplementedError("only can sampling for mf-type model")
        self.path = self._resolve_dir(args)
        self.expo_popularity = None
        self._load()
        self.users = range(self.n_users)
        self.items = range(self.n_items)
        self.train_user_list = CsrDict(self.train_indptr, self.train_items)
        self.train_user_list_time = CsrDict(self.train_indptr, self.train_times) if self.train_times is not None else None
        self.valid_user_list = CsrDict(self.valid_indptr, self.valid_items, self.valid_user_order)
        self.test_user_list = CsrDict(self.test_indptr, self.test_items, self.test_user_order)
        self.valid_users = self.valid_user_list.keys()
        self.test_users = set(self.test_user_list.keys())
        print('n_items:', self.n_items, 'n_users:', self.n_users)
        print("sparsity:", 1.0 * self.n_train / self.n_items / self.n_users)

    @staticmethod
    def _resolve_dir(args):
        first = './data/{}/'.format(args.dataset)
        if os.path.isdir(first):
            return first
        alt = os.path.join(getattr(args, "data_path", "./data/"), args.dataset) + "/"
        if os.path.isdir(alt):
            return alt
        raise FileNotFoundError("no dataset directory: tried %s and %s" % (first, alt))

    # ---- cache ----
    def _cache_path(self):
        return os.path.join(self.path, "pda_cache_%s.npz" % self.kind)

    _FIELDS = ("n_users", "n_items", "n_train", "n_valid", "n_test", "train_indptr", "train_items", "train_times",
               "unique_times", "valid_indptr", "valid_items", "valid_user_order", "test_indptr", "test_items",
               "test_user_order")

    def _load(self):
        cache, srcs = self._cache_path(), [os.path.join(self.path, f) for f in self._source_files()]
        have_src = all(os.path.exists(s) for s in srcs)
        if os.path.exists(cache) and (not have_src or os.path.getmtime(cache) >= max(os.path.getmtime(s) for s in srcs)):
            z = np.load(cache)
            if int(z["version"]) == CACHE_VERSION:
                for k in self._FIELDS:
                    v = z[k]
                    setattr(self, k, int(v) if v.ndim == 0 else v)
                if self.train_times.size == 0 and self.kind == "Data":
                    self.train_times = None
                self.unique_times = [int(t) for t in self.unique_times]
                return
        if not have_src:
            raise FileNotFoundError("missing interaction files under %s: %s" % (self.path, self._source_files()))
        self._parse()
        try:
            d = {k: getattr(self, k) for k in self._FIELDS}
            if d["train_times"] is None:
                d["train_times"] = np.zeros(0, dtype=np.uint8)
            d["unique_times"] = np.asarray(d["unique_times"], dtype=np.int32)
            np.savez_compressed(cache, version=CACHE_VERSION, **d)
        except OSError:
            pass   # read-only data directory: parse again next time

    def _load_eval_lists(self):
        vu, vc, vi = _parse_user_lists(os.path.join(self.path, "valid.txt"))
        tu, tc, ti = _parse_user_lists(os.path.join(self.path, "test.txt"))
        self.n_valid, self.n_test = int(vc.sum()), int(tc.sum())
        return (vu, vc, vi), (tu, tc, ti)

    def _finish_eval(self, v, t):
        (vu, vc, vi), (tu, tc, ti) = v, t
        self.valid_indptr, self.valid_items, _ = _csr_from_lists(self.n_users, vu, vc, vi, sort_rows=False)
        self.test_indptr, self.test_items, _ = _csr_from_lists(self.n_users, tu, tc, ti, sort_rows=False)
        self.valid_user_order, self.test_user_order = vu.astype(np.int64), tu.astype(np.int64)

    def add_expo_popularity(self, popularity):
        self.expo_popularity = popularity


class Data(_Interactions):
    """train.txt based set of BPRMF (load_data.py:24-110)."""
    kind = "Data"

    def _source_files(self):
        return ["train.txt", "valid.txt", "test.txt"]

    def _parse(self):
        u, c, it = _parse_user_lists(os.path.join(self.path, "train.txt"))
        v, t = self._load_eval_lists()
        mx_u = max([int(u.max())] + [int(x[0].max()) for x in (v, t) if len(x[0])])
        mx_i = max([int(it.max())] + [int(x[2].max()) for x in (v, t) if len(x[2])])
        self.n_users, self.n_items, self.n_train = mx_u + 1, mx_i + 1, int(c.sum())
        self.train_indptr, self.train_items, _ = _csr_from_lists(self.n_users, u, c, it, sort_rows=True)
        self.train_times, self.unique_times = None, []
        print(self.n_train, self.n_valid, self.n_test)
        self._finish_eval(v, t)


class Data2(_Interactions):
    """train_with_time.txt based set of PD / PDA / BPR(t)-pop (load_data.py:617-710)."""
    kind = "Data2"

    def _source_files(self):
        return ["train_with_time.txt", "valid.txt", "test.txt"]

    def _parse(self):
        import pandas as pd
        df = pd.read_csv(os.path.join(self.path, "train_with_time.txt"), header=None, sep=" ")
        uid = df[0].to_numpy().astype(np.int64)
        iid = df[1].to_numpy().astype(np.int64)
        tim = df[2].to_numpy().astype(np.int64)
        _, first = np.unique(tim, return_index=True)
        self.unique_times = [int(x) for x in tim[np.sort(first)]]       # order of appearance, like Series.unique()
        print("time slot unique in train:", np.array(self.unique_times))
        if len(self.unique_times) < 2:
            raise RuntimeWarning("there only one time slot for train...., this may cause our method not work")
        if tim.min() < 0 or tim.max() > 255:
            raise ValueError("stage labels must fit uint8")
        v, t = self._load_eval_lists()
        mx_u = max([int(uid.max())] + [int(x[0].max()) for x in (v, t) if len(x[0])])
        mx_i = max([int(iid.max())] + [int(x[2].max()) for x in (v, t) if len(x[2])])
        self.n_users, self.n_items, self.n_train = mx_u + 1, mx_i + 1, int(len(uid))
        order = np.lexsort((np.arange(len(uid)), iid, uid))
        deg = np.bincount(uid, minlength=self.n_users).astype(np.int64)
        self.train_indptr = np.zeros(self.n_users + 1, dtype=np.int64)
        np.cumsum(deg, out=self.train_indptr[1:])
        self.train_items = iid[order].astype(np.int32)
        self.train_times = tim[order].astype(np.uint8)
        print(self.n_train, self.n_valid, self.n_test)
        self._finish_eval(v, t)
