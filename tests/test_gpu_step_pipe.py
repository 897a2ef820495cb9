"""The bulk-copy pipelined step kernel (pda_step_pipe.cu, d = 128) against the register-gather kernel (pda_train.cu) and
the CPU oracle: same operations in the same order => the same bits, for every ring depth / CTA shape / L2-policy
setting, ragged batch sizes, long replay lags, out-of-range Adam operands (generic path) and never-touched rows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pda():
    import pda_b200
    assert pda_b200.load().pda_device_count() >= 1
    return pda_b200


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


def _state(m):
    return {k: m.get_table(k) for k in ("user_embedding", "item_embedding", "user_m", "user_v", "item_m", "item_v")}


def _distinct_batch(rng, n_users, n_items, B):
    users = rng.permutation(n_users)[:B].astype(np.int32)
    perm = rng.permutation(n_items)          # distinct items inside a batch: no atomic-order noise -> bit-exact
    return users, perm[:B].astype(np.int32), perm[B:2 * B].astype(np.int32), rng.random(B).astype(np.float32), rng.random(B).astype(np.float32)


@pytest.mark.parametrize("train", ["s_condition", "normal"])
@pytest.mark.parametrize("adam", ["lazy"])
@pytest.mark.parametrize("knobs", [{}, {"PDA_STEP_PIPE_D": "2"}, {"PDA_STEP_PIPE_D": "4"}, {"PDA_STEP_PIPE_V": "0"},
                                   {"PDA_STEP_PIPE_NW": "4"}, {"PDA_STEP_PIPE_HINTS": "3"},
                                   {"PDA_STEP_PIPE_NW": "4", "PDA_STEP_PIPE_V": "0"}, {"hot": "28"}, {"hot": "28", "PDA_STEP_PIPE_NW": "4"},
                                   {"hot": "5", "PDA_STEP_PIPE_D": "2"}])
def test_pipe_kernel_bit_identical_to_register_kernel(pda, monkeypatch, train, adam, knobs):
    """The fused user-row Adam (replay + update in the step kernel; the dense-table variant of the pipelined kernel needs
    the device sampler's distinct-users guarantee and is covered by the oracle test below).  Host batches with distinct users and distinct items -> every table and Adam slot bit-identical after every step,
    ragged sizes (1, 31, 33, 1000, 4097) included; users come back after long lags."""
    if knobs and train == "normal" and knobs != {"PDA_STEP_PIPE_V": "0"}:
        pytest.skip("knob sweep on the main variant only")
    rng = np.random.default_rng(5)
    n_users, n_items, d = 9000, 9000, 128
    U = rng.normal(0, 0.1, (n_users, d)).astype(np.float32)
    I = rng.normal(0, 0.1, (n_items, d)).astype(np.float32)
    ms = {}
    for name in ("pipe", "reg"):
        m = pda.PDAModel(n_users, n_items, d, train=train, batch_size=4097, lr=1e-2, regs=1e-3, init=False, max_batch=4097)
        m.set_table("user_embedding", U); m.set_table("item_embedding", I)
        m.set_adam_mode(adam)
        ms[name] = m
    knobs = dict(knobs)
    if "hot" in knobs:   # popular-item rows pre-summed in shared memory: an item occurs once per batch here -> still the same bits
        ms["pipe"].set_hot_items(rng.permutation(n_items)[:int(knobs.pop("hot"))])
    sizes = [1000, 1, 31, 33, 4097, 1000, 2, 4096, 64] + [777] * 12
    for step, B in enumerate(sizes):
        batch = _distinct_batch(rng, n_users, n_items, B)
        if train == "normal":
            batch = batch[:3]
        losses = {}
        for name, m in ms.items():
            monkeypatch.setenv("PDA_STEP_PIPE", "1" if name == "pipe" else "0")
            for k, v in knobs.items():
                monkeypatch.setenv(k, v)
            # the distinct-users check of the host path picks the fused kernel in lazy mode (uniq_users)
            losses[name] = m.train_step(*batch)
        # loss scalars: fp32 partial sums of squares + fp64 atomics across warps -> not bit-reproducible by design (1e-5 contract)
        assert np.allclose(losses["pipe"], losses["reg"], rtol=1e-6, atol=0), (step, B, losses)
        if step in (0, 4, 8, len(sizes) - 1):
            a, b = _state(ms["pipe"]), _state(ms["reg"])
            for k in a:
                assert np.array_equal(bits(a[k]), bits(b[k])), (step, B, k)
    if adam == "lazy":
        assert ms["pipe"].adam_stats()[1] > 0          # rows did replay skipped steps inside the pipelined kernel
    for m in ms.values():
        m.close()


@pytest.mark.parametrize("adam", ["lazy", "dense"])
def test_pipe_kernel_matches_oracle_with_duplicate_items_and_sampler(pda, c_oracle, adam):
    """lazy = fused user-row Adam in the kernel, dense = user gradient stored + dense sweep (both pipelined variants);
    device-sampled batches (items repeat -> fp32 atomics order) through the pipelined fused kernel vs the C oracle's dense
    TF1 Adam: losses 1e-5, tables 1e-4 of scale, never-sampled user rows untouched."""
    from helpers import pop_table, synth_interactions
    from oracle import pda_oracle as po
    n_users, n_items, T, B, d = 7000, 3000, 9, 1000, 128
    uid, iid, t = synth_interactions(n_users, n_items, 8, T, seed=21)
    indptr, items, times = po.build_csr(n_users, uid, iid, t)
    P = po.train_pop_matrix(pop_table(n_items, T, 2), 0.16)
    m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=B, lr=1e-2, regs=1e-3, seed=2021, max_batch=B)
    m.set_adam_mode(adam)
    m.set_train_csr(indptr, items, times, unique_times=np.arange(T))
    m.set_train_pop(P)
    ref = c_oracle.CModel(m.get_table("user_embedding"), m.get_table("item_embedding"), 1e-2, 1e-3, B, "s_condition")
    active = np.nonzero(np.diff(indptr) > 0)[0]
    for s in range(25):
        m.train_sampled(2020, 0, s, 1, B)
        got = m.read_loss()
        b = c_oracle.sample_batch(2020, 0, s, B, active, indptr, items, times, n_items, np.arange(T), P)
        want = ref.train_step(b["users"], b["pos"], b["neg"], b["pos_pop"], b["neg_pop"])
        assert np.allclose(got, want, rtol=1e-5, atol=0), (s, got, want)
    m.train_sampled(2020, 0, 25, 15, B)          # the overlapped-sampler path, 15 steps in one call
    for s in range(25, 40):
        b = c_oracle.sample_batch(2020, 0, s, B, active, indptr, items, times, n_items, np.arange(T), P)
        want = ref.train_step(b["users"], b["pos"], b["neg"], b["pos_pop"], b["neg_pop"])
    assert np.allclose(m.read_loss(), want, rtol=1e-5, atol=0)
    for name, r in (("user_embedding", ref.U), ("item_embedding", ref.I), ("user_m", ref.mU), ("user_v", ref.vU)):
        g = m.get_table(name)
        assert np.abs(g - r).max() <= 1e-4 * np.abs(r).max(), name
    m.close()


@pytest.mark.parametrize("variant", ["1", "0"])
def test_pipe_kernel_generic_path_for_out_of_range_operands(pda, monkeypatch, variant):
    """Adam slots outside the guarded ranges (tiny / huge / zero m, v; some lanes zero) take the generic
    __fsqrt_rn / __fdiv_rn path inside the pipelined kernel: still bit-identical to the register kernel and to the
    dense sweep."""
    rng = np.random.default_rng(9)
    n_users, n_items, d, B = 4000, 5000, 128, 512
    U = rng.normal(0, 0.1, (n_users, d)).astype(np.float32)
    I = rng.normal(0, 0.1, (n_items, d)).astype(np.float32)
    # |m| from denormal to 1e7 per row, v = m^2 * 10^k (k in [-2, 2]): sqrt / quotient operands far outside the guarded
    # ranges on many rows (v underflows to 0 or denormals, v > 2^40, |lr m| < 2^-100), the update itself stays O(lr)
    m64 = rng.normal(0, 1, (n_users, d)) * 10.0 ** rng.integers(-42, 8, (n_users, 1))
    mU = m64.astype(np.float32)
    with np.errstate(under="ignore", over="ignore"):
        vU = ((m64 * rng.uniform(0.5, 2.0, (n_users, d))) ** 2 * 10.0 ** rng.integers(-2, 3, (n_users, 1))).astype(np.float32)
    mU[::7, 4:8] = 0.0; vU[::7, 4:8] = 0.0              # one lane's float4 all zero
    mU[::11] = 0.0; vU[::11] = 0.0                      # "never touched" rows
    ms = {}
    for name, adam in (("pipe", "lazy"), ("reg", "lazy"), ("dense", "dense")):
        m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=B, lr=1e-2, regs=1e-3, init=False)
        m.set_table("user_embedding", U); m.set_table("item_embedding", I)
        m.set_table("user_m", mU); m.set_table("user_v", vU)
        m.set_adam_mode(adam)
        ms[name] = m
    for step in range(14):
        batch = _distinct_batch(rng, n_users, n_items, B)
        out = {}
        for name, m in ms.items():
            monkeypatch.setenv("PDA_STEP_PIPE", "0" if name == "reg" else "1")
            monkeypatch.setenv("PDA_STEP_PIPE_V", variant)      # 1: warp-voted guard + negated-v replay, 0: per-lane guards
            out[name] = m.train_step(*batch)
        assert np.allclose(out["pipe"], out["reg"], rtol=1e-6, atol=0, equal_nan=True), (step, out)
        assert np.allclose(out["pipe"], out["dense"], rtol=1e-6, atol=0, equal_nan=True), (step, out)
    a, b, c = _state(ms["pipe"]), _state(ms["reg"]), _state(ms["dense"])
    for k in a:
        assert np.array_equal(bits(a[k]), bits(b[k])), k
        assert np.array_equal(bits(a[k]), bits(c[k])), k
    for m in ms.values():
        m.close()


@pytest.mark.parametrize("train", ["s_condition", "normal"])
def test_popular_item_rows_summed_in_shared_memory(pda, c_oracle, train):
    """A Zipf head on the positive items (half of the batch on 28 items, one item on 10 % of it): the per-CTA shared-memory
    sums of the popular rows (pda_set_hot_items) give the item gradient of the plain red.global.add path and of the C
    oracle to 1e-5 of scale (item_m after the first Adam step = 0.1 x gradient), for hot lists of 0 / 5 / 28 items;
    ids outside the table and repeated ids are rejected."""
    rng = np.random.default_rng(31)
    n_users, n_items, d, B = 6000, 2000, 128, 4096
    U = rng.normal(0, 0.1, (n_users, d)).astype(np.float32)
    I = rng.normal(0, 0.1, (n_items, d)).astype(np.float32)
    head = rng.permutation(n_items)[:28].astype(np.int32)
    users = rng.permutation(n_users)[:B].astype(np.int32)
    pos = rng.integers(0, n_items, B).astype(np.int32)
    sel = rng.random(B)
    pos[sel < 0.5] = head[rng.integers(0, 28, int((sel < 0.5).sum()))]
    pos[sel < 0.1] = head[0]
    neg = rng.integers(0, n_items, B).astype(np.int32)
    pp, pn = rng.random(B).astype(np.float32), rng.random(B).astype(np.float32)
    batch = (users, pos, neg, pp, pn) if train == "s_condition" else (users, pos, neg)
    ref = c_oracle.CModel(U, I, 1e-2, 1e-3, B, train)
    want_loss = ref.train_step(users, pos, neg, pp if train == "s_condition" else None, pn if train == "s_condition" else None)
    scale = np.abs(ref.mI).max()
    got = {}
    for n_hot in (0, 5, 28):
        m = pda.PDAModel(n_users, n_items, d, train=train, batch_size=B, lr=1e-2, regs=1e-3, init=False, max_batch=B)
        m.set_table("user_embedding", U); m.set_table("item_embedding", I)
        m.set_adam_mode("lazy")
        m.set_hot_items(head[:n_hot])
        loss = m.train_step(*batch)
        assert np.allclose(loss, want_loss, rtol=1e-5, atol=0), (n_hot, loss, want_loss)
        got[n_hot] = m.get_table("item_m")
        assert np.abs(got[n_hot] - ref.mI).max() <= 1e-5 * scale, n_hot
        assert np.array_equal(bits(m.get_table("user_embedding")), bits(ref.U)), n_hot      # the user rows are not touched by it
        if n_hot == 28:
            with pytest.raises(pda.PdaError, match="outside"):
                m.set_hot_items([3, n_items])
            with pytest.raises(pda.PdaError, match="twice"):
                m.set_hot_items([3, 4, 3])
        m.close()
    assert np.abs(got[28] - got[0]).max() <= 1e-5 * scale        # only the summation order of the popular rows differs (fp32 sums of up to ~400 rows)


def test_host_batches_with_bad_ids_are_rejected(pda):
    """ids outside their table: PDA_ERR_ARG from every host entry point (the reference's embedding_lookup raises)."""
    m = pda.PDAModel(100, 50, 16, train="normal", batch_size=8)
    ok = np.arange(8, dtype=np.int32)
    for users, pos, neg in ((ok + 100, ok, ok), (ok, ok + 50, ok), (ok, ok, -1 - ok)):
        with pytest.raises(pda.PdaError, match="outside"):
            m.train_step(users, pos, neg)
        with pytest.raises(pda.PdaError, match="outside"):
            m.gradients(users, pos, neg)
    m.train_step(ok, ok, ok + 1)       # the model is still usable
    with pytest.raises(pda.PdaError, match="outside"):
        m.do_recommendation(np.array([5, 100], dtype=np.int32), None, "main_branch", K=5, mask=False)
    # a stage label without a column in the popularity table
    m2 = pda.PDAModel(10, 20, 16, train="s_condition", batch_size=4)
    indptr = np.arange(0, 11, dtype=np.int64)
    m2.set_train_csr(indptr, np.arange(10, dtype=np.int32), np.full(10, 5, dtype=np.uint8), unique_times=np.arange(6))
    with pytest.raises(pda.PdaError, match="stage label"):
        m2.set_train_pop(np.ones((20, 3), dtype=np.float32))
    m.close(); m2.close()
