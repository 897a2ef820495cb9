import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pda_b200 as pda

def bits(a): return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)
rng = np.random.default_rng(5)
n_users, n_items, d = 9000, 9000, 128
U = rng.normal(0, 0.1, (n_users, d)).astype(np.float32)
I = rng.normal(0, 0.1, (n_items, d)).astype(np.float32)
ms = {}
for name in ("pipe", "reg"):
    m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=4097, lr=1e-2, regs=1e-3, init=False, max_batch=4097)
    m.set_table("user_embedding", U); m.set_table("item_embedding", I)
    m.set_adam_mode("lazy")
    ms[name] = m
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1000, 1, 31, 33, 4097, 1000]
last_seen = np.full(n_users, -1)
for step, B in enumerate(sizes):
    users = rng.permutation(n_users)[:B].astype(np.int32)
    perm = rng.permutation(n_items)
    batch = (users, perm[:B].astype(np.int32), perm[B:2 * B].astype(np.int32), rng.random(B).astype(np.float32), rng.random(B).astype(np.float32))
    for name, m in ms.items():
        os.environ["PDA_STEP_PIPE"] = "1" if name == "pipe" else "0"
        l = m.train_step(*batch)
        print(step, B, name, l)
    for k in ("user_embedding", "user_m", "user_v", "item_embedding", "item_m", "item_v"):
        a, b = ms["pipe"].get_table(k), ms["reg"].get_table(k)
        bad = np.nonzero((bits(a) != bits(b)).any(axis=1))[0]
        if len(bad):
            inb = np.isin(bad, users if k.startswith("user") else np.concatenate([batch[1], batch[2]]))
            r = bad[0]
            cols = np.nonzero(bits(a[r]) != bits(b[r]))[0]
            print("  step", step, k, "rows differing", len(bad), "in batch", int(inb.sum()), "first row", r, "cols", cols[:8], len(cols),
                  "lag", (step - last_seen[r]) if k.startswith("user") else None,
                  "pos in batch", int(np.nonzero(users == r)[0][0]) if k.startswith("user") and r in users else None,
                  "a", a[r, cols[:3]], "b", b[r, cols[:3]], "maxabs", np.abs(a[bad] - b[bad]).max())
            if k.startswith("user"):
                pos_in = [int(np.nonzero(users == x)[0][0]) for x in bad[:40] if x in users]
                print("   positions of differing rows in the batch (first 40):", pos_in, " lags:", [int(step - last_seen[x]) for x in bad[:40]])
    last_seen[users] = step
