"""Stages the shipped Douban set for GPU runs: parses data/douban/douban.zip of the reference ONCE (here, where
/root/reference exists) and leaves only the parsed caches under ./data/douban/ (git-ignored, travels with gpurun):
  pda_cache_Data2.npz  (train_with_time.txt + valid/test)   pda_cache_Data.npz (train.txt + valid/test)
  pda_cache_pop.npy    (item_pop_seq_ori2.txt, float64 [26047, 10])
No reference file is copied into the repo; the loaders (pda_b200/data.py, popularity.py) read these caches
when the text files are absent."""
import argparse
import os
import shutil
import sys
import tempfile
import zipfile
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--zip", default="/root/reference/data/douban/douban.zip")
    ap.add_argument("--out", default=os.path.join(ROOT, "data", "douban"))
    a = ap.parse_args()
    from pda_b200 import data as D
    tmp = tempfile.mkdtemp(prefix="pda_douban_")
    try:
        dst = os.path.join(tmp, "data", "douban")
        os.makedirs(dst)
        z = zipfile.ZipFile(a.zip)
        for n in z.namelist():
            base = os.path.basename(n)
            if base in ("train_with_time.txt", "train.txt", "valid.txt", "test.txt", "item_pop_seq_ori2.txt"):
                with open(os.path.join(dst, base), "wb") as f:
                    f.write(z.read(n))
        cwd = os.getcwd()
        os.chdir(tmp)
        args = SimpleNamespace(dataset="douban", batch_size=2048, model="mf", data_path="./data/")
        d2 = D.Data2(args)
        d1 = D.Data(args)
        raw = np.loadtxt(os.path.join(dst, "item_pop_seq_ori2.txt"), dtype=np.float64, ndmin=2)
        os.chdir(cwd)
        os.makedirs(a.out, exist_ok=True)
        for f in ("pda_cache_Data2.npz", "pda_cache_Data.npz"):
            shutil.copy(os.path.join(dst, f), os.path.join(a.out, f))
        np.save(os.path.join(a.out, "pda_cache_pop.npy"), raw[:, 1:])
        print("Data2:", d2.n_users, d2.n_items, d2.n_train, d2.n_valid, d2.n_test, "stages", d2.unique_times)
        print("Data :", d1.n_users, d1.n_items, d1.n_train, d1.n_valid, d1.n_test)
        print("item id column is 0..n-1 in order:", bool((raw[:, 0] == np.arange(len(raw))).all()))
        for f in os.listdir(a.out):
            print(f, os.path.getsize(os.path.join(a.out, f)) >> 10, "KiB")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
