"""GPU parity of the training path (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest

from helpers import pop_table, synth_interactions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pda():
    import pda_b200
    lib = pda_b200.load()
    assert lib.pda_device_count() >= 1, "no CUDA device visible: GPU tests cannot run"
    return pda_b200


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


@pytest.mark.parametrize("rows,d", [(1000, 64), (333, 128), (77, 20)])
def test_xavier_init_bit_exact(pda, c_oracle, rows, d):
    m = pda.PDAModel(rows, rows + 5, d, seed=2021)
    U = m.get_table("user_embedding")
    I = m.get_table("item_embedding")
    assert np.array_equal(bits(U), bits(c_oracle.xavier_init(rows, d, 2021, 0)))
    assert np.array_equal(bits(I), bits(c_oracle.xavier_init(rows + 5, d, 2021, 1)))
    a = np.sqrt(6.0 / (rows + d))
    assert np.abs(U).max() <= a and abs(U.mean()) < a / 10
    m.close()


@pytest.mark.parametrize("B,empty", [(256, 0.0), (512, 0.1), (5000, 0.05)])
def test_sampler_bit_exact(pda, c_oracle, B, empty):
    from oracle import pda_oracle as po
    n_users, n_items, T = 3000, 800, 9
    uid, iid, t = synth_interactions(n_users, n_items, 12, T, seed=5, empty_frac=empty)
    indptr, items, times = po.build_csr(n_users, uid, iid, t)
    P = po.train_pop_matrix(pop_table(n_items, T, 3), 0.22)
    m = pda.PDAModel(n_users, n_items, 16, train="s_condition", batch_size=B, max_batch=B)
    m.set_train_csr(indptr, items, times, unique_times=np.arange(T))
    m.set_train_pop(P)
    active = np.nonzero(np.diff(indptr) > 0)[0]
    for epoch, step in [(0, 0), (3, 17), (1, 4000)]:
        got = m.sample_batch(2020, epoch, step, B)
        ref = c_oracle.sample_batch(2020, epoch, step, B, active, indptr, items, times, n_items, np.arange(T), P)
        for k in ("users", "pos", "neg", "time"):
            assert np.array_equal(got[k], ref[k]), k
        assert np.array_equal(bits(got["pos_pop"]), bits(ref["pos_pop"]))
        assert np.array_equal(bits(got["neg_pop"]), bits(ref["neg_pop"]))
        if B <= len(active):
            assert len(np.unique(got["users"])) == B      # rd.sample: distinct users
        # negatives never in the user's train list (train_new_api.py:402-406)
        for u, n in zip(got["users"][:200], got["neg"][:200]):
            assert n not in items[indptr[u]:indptr[u + 1]]
    m.close()


def _random_batch(rng, n_users, n_items, B, unique_users=True, unique_items=False):
    users = rng.permutation(n_users)[:B] if unique_users else rng.integers(0, n_users, B)
    if unique_items:
        it = rng.permutation(n_items)[:2 * B]
        pos, neg = it[:B], it[B:]
    else:
        pos = (rng.random(B) ** 3 * n_items).astype(np.int64)   # popular items repeat
        neg = rng.integers(0, n_items, B)
    pp = rng.random(B).astype(np.float32)
    npop = rng.random(B).astype(np.float32)
    pp[rng.random(B) < 0.1] = 0.0
    return users.astype(np.int32), pos.astype(np.int32), neg.astype(np.int32), pp, npop


@pytest.mark.parametrize("d", [8, 20, 32, 64, 128, 256, 320])
@pytest.mark.parametrize("mode", ["normal", "s_condition"])
def test_gradients_and_loss(pda, c_oracle, d, mode):
    rng = np.random.default_rng(d)
    n_users, n_items, B = 700, 300, 512
    U = (rng.normal(0, 0.6, (n_users, d)) / np.sqrt(d) * 4).astype(np.float32)
    I = (rng.normal(0, 0.6, (n_items, d)) / np.sqrt(d) * 4).astype(np.float32)
    users, pos, neg, pp, npop = _random_batch(rng, n_users, n_items, B)
    m = pda.PDAModel(n_users, n_items, d, train=mode, batch_size=2048, lr=1e-2, regs=1e-3, max_batch=B, init=False)
    m.set_table("user_embedding", U)
    m.set_table("item_embedding", I)
    gU, gI, loss3 = m.gradients(users, pos, neg, pp, npop)
    l3, rU, rP, rN = c_oracle.forward_backward(U, I, users, pos, neg, 1e-3, 2048, mode, pp, npop)
    # user rows: one term each -> bit-exact
    refU = np.zeros_like(U)
    refU[users] = rU
    assert np.array_equal(bits(gU), bits(refU))
    # item rows: duplicates are summed by L2 atomics (order free) -> 1e-5 relative (north_star tolerance)
    refI = np.zeros((n_items, d), dtype=np.float64)
    np.add.at(refI, pos, rP.astype(np.float64))
    np.add.at(refI, neg, rN.astype(np.float64))
    scale = np.abs(refI).max()
    assert np.abs(gI - refI).max() <= 1e-5 * scale
    cnt = np.bincount(np.concatenate([pos, neg]), minlength=n_items)
    single = cnt == 1
    ref_single = np.zeros((n_items, d), dtype=np.float32)
    ref_single[pos] = rP
    ref_single[neg] = rN
    assert np.array_equal(bits(gI[single]), bits(ref_single[single]))      # no duplicate -> bit-exact
    assert np.allclose(loss3, l3, rtol=1e-5, atol=0)
    m.close()


@pytest.mark.parametrize("mode,d", [("s_condition", 64), ("normal", 64), ("s_condition", 128)])
def test_train_steps_bit_exact_without_duplicates(pda, c_oracle, mode, d):
    """With distinct items in every batch no fp32 atomic ever meets another: the whole trajectory
    (gather, loss chain, gradients, TF1 dense Adam incl. untouched rows) must match the oracle bit for bit."""
    rng = np.random.default_rng(7)
    n_users, n_items, B = 900, 1200, 256
    U = (rng.normal(0, 0.3, (n_users, d))).astype(np.float32)
    I = (rng.normal(0, 0.3, (n_items, d))).astype(np.float32)
    m = pda.PDAModel(n_users, n_items, d, train=mode, batch_size=B, lr=1e-2, regs=1e-3, init=False)
    m.set_table("user_embedding", U)
    m.set_table("item_embedding", I)
    ref = c_oracle.CModel(U, I, 1e-2, 1e-3, B, mode)
    for step in range(6):
        users, pos, neg, pp, npop = _random_batch(rng, n_users, n_items, B, unique_items=True)
        got = m.train_step(users, pos, neg, pp, npop)
        want = ref.train_step(users, pos, neg, pp, npop)
        assert np.allclose(got, want, rtol=1e-5, atol=0), (step, got, want)
    assert np.array_equal(bits(m.get_table("user_embedding")), bits(ref.U))
    assert np.array_equal(bits(m.get_table("item_embedding")), bits(ref.I))
    assert np.array_equal(bits(m.get_table("item_m")), bits(ref.mI))
    assert np.array_equal(bits(m.get_table("user_v")), bits(ref.vU))
    assert np.array_equal(bits(m.get_adam_powers()), bits(ref.pw))
    # TF1 semantics: rows never sampled still moved (dense m/v decay + update), SURVEY A.4
    m.close()


def test_train_steps_with_duplicates_tolerance(pda, c_oracle):
    rng = np.random.default_rng(11)
    n_users, n_items, B, d = 3000, 500, 1024, 64
    m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=B, lr=1e-2, regs=1e-3, seed=2021)
    U0, I0 = m.get_table("user_embedding"), m.get_table("item_embedding")
    ref = c_oracle.CModel(U0, I0, 1e-2, 1e-3, B, "s_condition")
    for step in range(20):
        users, pos, neg, pp, npop = _random_batch(rng, n_users, n_items, B)
        got = m.train_step(users, pos, neg, pp, npop)
        want = ref.train_step(users, pos, neg, pp, npop)
        assert np.allclose(got, want, rtol=1e-5, atol=0), (step, got, want)
    # per-step quantities hold 1e-5; the 20-step Adam trajectory amplifies the atomics' last-bit
    # differences through m / (sqrt(v) + eps) (SURVEY 7 "trajectory divergence") -> 1e-4 of the table scale
    for name, r in (("user_embedding", ref.U), ("item_embedding", ref.I)):
        g = m.get_table(name)
        assert np.abs(g - r).max() <= 1e-4 * np.abs(r).max(), name
    untouched = np.setdiff1d(np.arange(n_items), np.concatenate([pos, neg]))
    assert len(untouched) and not np.array_equal(m.get_table("item_embedding")[untouched], I0[untouched])
    m.close()


def test_sampled_training_matches_oracle_pipeline(pda, c_oracle):
    """device sampler -> fused step -> Adam, n steps enqueued back to back, vs the oracle doing the same."""
    from oracle import pda_oracle as po
    n_users, n_items, T, B, d = 2500, 600, 9, 512, 64
    uid, iid, t = synth_interactions(n_users, n_items, 10, T, seed=9)
    indptr, items, times = po.build_csr(n_users, uid, iid, t)
    P = po.train_pop_matrix(pop_table(n_items, T, 4), 0.16)
    m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=B, lr=1e-2, regs=1e-3, seed=2021)
    m.set_train_csr(indptr, items, times, unique_times=np.arange(T))
    m.set_train_pop(P)
    ref = c_oracle.CModel(m.get_table("user_embedding"), m.get_table("item_embedding"), 1e-2, 1e-3, B, "s_condition")
    active = np.nonzero(np.diff(indptr) > 0)[0]
    n_steps = 12
    m.train_sampled(2020, 0, 0, n_steps, B)
    got = m.read_loss()
    for s in range(n_steps):
        b = c_oracle.sample_batch(2020, 0, s, B, active, indptr, items, times, n_items, np.arange(T), P)
        want = ref.train_step(b["users"], b["pos"], b["neg"], b["pos_pop"], b["neg_pop"])
    assert np.allclose(got, want, rtol=1e-5, atol=0)
    for name, r in (("user_embedding", ref.U), ("item_embedding", ref.I)):
        g = m.get_table(name)
        assert np.abs(g - r).max() <= 1e-4 * np.abs(r).max(), name
    m.close()


def test_error_paths(pda):
    with pytest.raises(pda.PdaError):
        pda.PDAModel(10, 10, 7)               # embed_size not a multiple of 4
    m = pda.PDAModel(50, 40, 8, train="s_condition", batch_size=16)
    with pytest.raises(pda.PdaError):
        m.sample_batch(1, 0, 0)               # no CSR yet
    with pytest.raises(pda.PdaError):
        m.train_step(np.arange(17), np.arange(17), np.arange(17), np.ones(17), np.ones(17))   # B > capacity
    with pytest.raises(pda.PdaError):
        m.set_train_csr(np.array([0] + [2] * 50), np.array([5, 3]))   # unsorted row
    m.close()


# ---------------------------------------------------------------------------------------------
# exact lazy replay of the TF1 dense Adam sweep (pda_adam_lazy.cu)
# ---------------------------------------------------------------------------------------------
def _all_state(m):
    return {k: m.get_table(k) for k in ("user_embedding", "item_embedding", "user_m", "user_v", "item_m", "item_v")}


@pytest.mark.parametrize("mode,d", [("s_condition", 64), ("normal", 128), ("s_condition", 20)])
def test_lazy_adam_bit_identical_to_dense(pda, mode, d):
    """Same batches through the dense sweep and through the lazy replay: every table and Adam slot must agree bit for
    bit at every read-out, including rows that were never sampled, rows sampled once and then left alone for many
    steps, and continuing after a read-out (which flushes)."""
    rng = np.random.default_rng(5)
    n_users, n_items, B = 3000, 2500, 128
    U = rng.normal(0, 0.3, (n_users, d)).astype(np.float32)
    I = rng.normal(0, 0.3, (n_items, d)).astype(np.float32)
    models = {}
    for am in ("dense", "lazy"):
        m = pda.PDAModel(n_users, n_items, d, train=mode, batch_size=B, lr=1e-2, regs=1e-3, init=False)
        m.set_adam_mode(am)
        m.set_table("user_embedding", U)
        m.set_table("item_embedding", I)
        models[am] = m
    for step in range(40):
        users, pos, neg, pp, npop = _random_batch(rng, n_users, n_items, B, unique_items=True)
        if step >= 25:          # a phase that only touches a few rows: everything else accumulates a long lag
            users = users % 200
            users = np.unique(users)
            k = len(users)
            pos, neg, pp, npop = pos[:k] % 300, 300 + neg[:k] % 300, pp[:k], npop[:k]
            _, fi = np.unique(pos, return_index=True)
            users, pos, neg, pp, npop = users[fi], pos[fi], neg[fi], pp[fi], npop[fi]
            _, fi = np.unique(neg, return_index=True)
            users, pos, neg, pp, npop = users[fi], pos[fi], neg[fi], pp[fi], npop[fi]
        la = models["lazy"].train_step(users, pos, neg, pp, npop)
        de = models["dense"].train_step(users, pos, neg, pp, npop)
        assert la == de, (step, la, de)
        if step in (0, 7, 24, 39):
            a, b = _all_state(models["lazy"]), _all_state(models["dense"])
            for k in a:
                assert np.array_equal(bits(a[k]), bits(b[k])), (step, k)
    assert np.array_equal(bits(models["lazy"].get_adam_powers()), bits(models["dense"].get_adam_powers()))
    for m in models.values():
        m.close()


def test_lazy_adam_mode_switches_and_eval_flush(pda, c_oracle):
    """dense -> lazy -> dense in one run equals the oracle's dense run; recommending in lazy mode sees current tables."""
    rng = np.random.default_rng(8)
    n_users, n_items, B, d = 1500, 5000, 128, 64
    U = rng.normal(0, 0.3, (n_users, d)).astype(np.float32)
    I = rng.normal(0, 0.3, (n_items, d)).astype(np.float32)
    m = pda.PDAModel(n_users, n_items, d, train="normal", batch_size=B, lr=1e-2, regs=1e-3, init=False)
    m.set_table("user_embedding", U); m.set_table("item_embedding", I)
    ref = c_oracle.CModel(U, I, 1e-2, 1e-3, B, "normal")
    for step in range(30):
        if step == 0:
            m.set_adam_mode("dense")
        if step == 10:
            m.set_adam_mode("lazy")
        if step == 22:
            m.set_adam_mode("dense")
        users, pos, neg, pp, npop = _random_batch(rng, n_users, n_items, B, unique_items=True)
        m.train_step(users, pos, neg)
        ref.train_step(users, pos, neg)
        if step == 18:      # lazy phase: scoring must flush first
            eu = np.arange(700, dtype=np.int32)
            for backend in ("exact", "tensor"):
                ids = m.do_recommendation(eu, None, "main_branch", K=20, mask=False, backend=backend)
                rid, _ = c_oracle.recommend(ref.U, ref.I, eu, "main_branch", 20)
                assert np.array_equal(ids, rid), backend
    assert np.array_equal(bits(m.get_table("user_embedding")), bits(ref.U))
    assert np.array_equal(bits(m.get_table("item_embedding")), bits(ref.I))
    assert np.array_equal(bits(m.get_table("item_v")), bits(ref.vI))
    m.close()


@pytest.mark.parametrize("d", [64, 128])
def test_fused_user_adam_in_step_kernel_matches_dense(pda, d):
    """Device-sampled batches (distinct users) in lazy mode take the fused path: the step kernel replays and updates the
    user rows itself.  Against the dense sweep on the same sampled batches: user rows with long lags, never-sampled rows,
    read-outs in the middle.  Items repeat inside a batch (fp32 atomics order, amplified by m / (sqrt(v) + eps) over the
    steps), hence 1e-4 of the table scale like the other multi-step trajectory checks."""
    from oracle import pda_oracle as po
    n_users, n_items, T, B = 6000, 3000, 9, 256
    uid, iid, t = synth_interactions(n_users, n_items, 8, T, seed=21)
    indptr, items, times = po.build_csr(n_users, uid, iid, t)
    P = po.train_pop_matrix(pop_table(n_items, T, 2), 0.16)
    models = {}
    for am in ("dense", "lazy"):
        m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=B, lr=1e-2, regs=1e-3, seed=2021)
        m.set_adam_mode(am)
        m.set_train_csr(indptr, items, times, unique_times=np.arange(T))
        m.set_train_pop(P)
        models[am] = m
    done = 0
    for n_steps in (1, 30, 3, 60):
        for m in models.values():
            m.train_sampled(2020, 0, done, n_steps, B)
        done += n_steps
        assert np.allclose(models["lazy"].read_loss(), models["dense"].read_loss(), rtol=1e-5, atol=0)
        a, b = _all_state(models["lazy"]), _all_state(models["dense"])
        for k in a:
            scale = np.abs(b[k]).max()
            assert np.abs(a[k] - b[k]).max() <= 1e-4 * scale, (done, k, np.abs(a[k] - b[k]).max(), scale)
    rows_updated, replayed = models["lazy"].adam_stats()
    assert replayed > 0          # users did come back after skipping steps, and were replayed in the step kernel
    # rows never sampled: untouched by both (m = v = 0 -> the dense sweep moves nothing either)
    U0 = po.xavier_init(n_users, d, 2021, 0)
    same = (models["dense"].get_table("user_m") == 0).all(axis=1)
    assert same.sum() > 0 and np.array_equal(models["lazy"].get_table("user_embedding")[same], U0[same])
    for m in models.values():
        m.close()


@pytest.mark.parametrize("adam", ["dense", "lazy"])
def test_pipelined_host_batches_equal_single_steps(pda, adam):
    """pda_train_steps_host (copies of batch k+1 overlapped with step k, two device slots) == n calls of
    pda_train_step_host, bit for bit -- distinct users (fused user-row path when lazy) and one batch with a repeated
    user in the middle (the device-side check must route that batch through the general path)."""
    rng = np.random.default_rng(17)
    n_users, n_items, d, B, n = 3000, 1100, 64, 512, 7
    U = rng.normal(0, 0.1, (n_users, d)).astype(np.float32)
    I = rng.normal(0, 0.1, (n_items, d)).astype(np.float32)
    ms = []
    for _ in range(2):
        m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=B, lr=1e-2, regs=1e-3, init=False)
        m.set_table("user_embedding", U); m.set_table("item_embedding", I)
        m.set_adam_mode(adam)
        ms.append(m)
    a, b = ms
    users = a.pinned_array((n, B), np.int32); pos = a.pinned_array((n, B), np.int32); neg = a.pinned_array((n, B), np.int32)
    pp = a.pinned_array((n, B), np.float32); pn = a.pinned_array((n, B), np.float32)
    for k in range(n):
        users[k] = rng.permutation(n_users)[:B]
        perm = rng.permutation(n_items)          # distinct items inside a batch: no atomic-order noise -> bit-exact
        pos[k], neg[k] = perm[:B], perm[B:2 * B]
    users[3, 5] = users[3, 4]                    # one batch with a repeated user
    pp[:] = rng.random((n, B)); pn[:] = rng.random((n, B))
    got = a.train_steps(users, pos, neg, pp, pn)
    want = np.asarray([b.train_step(users[k], pos[k], neg[k], pp[k], pn[k]) for k in range(n)], dtype=np.float32)
    assert np.array_equal(bits(got), bits(want))
    for t in ("user_embedding", "item_embedding"):
        assert np.array_equal(bits(a.get_table(t)), bits(b.get_table(t))), t
    with pytest.raises(pda.PdaError, match="pinned"):
        a.train_steps(np.array(users), pos, neg, pp, pn)
    a.close(); b.close()


def test_cuda_path_reproduces_the_committed_step_and_eval_golden(pda):
    """tests/golden/oracle_step_eval.npz: 3 PD steps on a 256 x 512 slice + PD / PDA top-50, committed
    vectors -- no oracle code runs in this test."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_step_eval.npz"))
    n_users, d = g["U0"].shape
    n_items = g["I0"].shape[0]
    B = len(g["b0_users"])
    for adam in ("dense", "lazy"):
        m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=B, lr=1e-2, regs=1e-3, init=False)
        m.set_table("user_embedding", g["U0"]); m.set_table("item_embedding", g["I0"])
        m.set_adam_mode(adam)
        for s in range(3):
            got = m.train_step(*[g[f"b{s}_{k}"] for k in ("users", "pos", "neg", "pos_pop", "neg_pop")])
            assert np.allclose(got, g["losses"][s], rtol=1e-5, atol=0), (adam, s)
        assert np.array_equal(bits(m.get_table("user_embedding")), bits(g["U3"])), adam
        assert np.array_equal(bits(m.get_table("item_embedding")), bits(g["I3"])), adam
        m.set_train_csr(g["mask_indptr"], g["mask_items"])
        for tag, rec, p in (("main", "main_branch", None), ("pda_last", "condition", g["pop_last"]),
                            ("pda_linear", "condition", g["pop_linear"])):
            ids, sc = m.do_recommendation(g["eval_users"], None, rec, pos_pop=p, K=int(g["K"]), backend="exact",
                                          return_scores=True)
            assert np.array_equal(ids, g[f"ids_{tag}"]), (adam, tag)
            assert np.array_equal(bits(sc), bits(g[f"scores_{tag}"])), (adam, tag)
        m.close()
