"""Edge cases of the hot path on the GPU (through the C ABI) against the oracle: tiny and ragged shapes, single rows,
users whose whole top of the ranking is masked, K at the limits, batches of one triple, repeated users."""
import numpy as np
import pytest

from helpers import pop_table, synth_interactions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pda():
    import pda_b200
    assert pda_b200.load().pda_device_count() >= 1, "no CUDA device visible: GPU tests cannot run"
    return pda_b200


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


@pytest.mark.parametrize("backend", ["exact", "tensor"])
@pytest.mark.parametrize("M,K", [(1, 1), (1, 50), (3, 128), (129, 7)])
def test_recommend_few_rows_and_k_limits(pda, c_oracle, backend, M, K):
    from oracle import pda_oracle as po
    rng = np.random.default_rng(M * 1000 + K)
    n_users, n_items, d = 300, 4099, 64          # n_items: not a multiple of any tile size
    U = rng.normal(0, 0.4, (n_users, d)).astype(np.float32)
    I = rng.normal(0, 0.4, (n_items, d)).astype(np.float32)
    uid, iid, t = synth_interactions(n_users, n_items, 30, 2, seed=5, empty_frac=0.2)
    indptr, items, _ = po.build_csr(n_users, uid, iid, t)
    pop = (rng.random(n_items) ** 2).astype(np.float32)
    m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=8, init=False)
    m.set_table("user_embedding", U); m.set_table("item_embedding", I)
    m.set_train_csr(indptr, items)
    users = rng.permutation(n_users)[:M].astype(np.int32)
    for rec_type in ("main_branch", "condition"):
        ids, sc = m.do_recommendation(users, None, rec_type, pos_pop=pop, K=K, backend=backend, return_scores=True)
        rid, rsc = c_oracle.recommend(U, I, users, rec_type, K, indptr, items, pop=pop)
        assert np.array_equal(ids, rid) and np.array_equal(bits(sc), bits(rsc)), (rec_type, backend)
    m.close()


def test_recommend_rows_with_almost_everything_masked(pda, c_oracle):
    """a user who interacted with all but 30 items: fewer than K unmasked items exist -> the tail of the top-K holds
    masked items at -inf in ascending id (tf.nn.top_k on ties), identically on both back ends."""
    from oracle import pda_oracle as po
    rng = np.random.default_rng(9)
    n_users, n_items, d, K = 64, 4200, 64, 50
    U = rng.normal(0, 0.4, (n_users, d)).astype(np.float32)
    I = rng.normal(0, 0.4, (n_items, d)).astype(np.float32)
    free = rng.permutation(n_items)[:30]
    heavy = np.setdiff1d(np.arange(n_items), free)
    uid = np.concatenate([np.zeros(len(heavy), np.int64), np.full(100, 5, np.int64)])
    iid = np.concatenate([heavy, rng.permutation(n_items)[:100]])
    indptr, items, _ = po.build_csr(n_users, uid, iid)
    m = pda.PDAModel(n_users, n_items, d, train="normal", batch_size=8, init=False)
    m.set_table("user_embedding", U); m.set_table("item_embedding", I)
    m.set_train_csr(indptr, items)
    users = np.array([0, 5, 7], dtype=np.int32)
    rid, rsc = c_oracle.recommend(U, I, users, "main_branch", K, indptr, items)
    assert np.isneginf(rsc[0, 30:]).all() and np.isfinite(rsc[0, :30]).all()
    for backend in ("exact", "tensor"):
        ids, sc = m.do_recommendation(users, None, "main_branch", K=K, backend=backend, return_scores=True)
        assert np.array_equal(ids, rid) and np.array_equal(bits(sc), bits(rsc)), backend
    m.close()


def test_recommend_all_scores_tied(pda, c_oracle):
    """zero tables: every score ties -> ids 0..K-1 in order (lower index first), PDA mode ranks by pop then id."""
    n_users, n_items, d, K = 10, 5000, 64, 20
    m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=8, init=False)
    m.set_table("user_embedding", np.zeros((n_users, d), np.float32))
    m.set_table("item_embedding", np.zeros((n_items, d), np.float32))
    users = np.arange(n_users, dtype=np.int32)
    pop = np.zeros(n_items, np.float32)
    pop[[4000, 17, 2500]] = [0.5, 0.5, 0.9]
    for backend in ("exact", "tensor"):
        ids = m.do_recommendation(users, None, "main_branch", K=K, mask=False, backend=backend)
        assert (ids == np.arange(K)[None, :]).all(), backend
        ids = m.do_recommendation(users, None, "condition", pos_pop=pop, K=K, mask=False, backend=backend)
        assert ids[0, :3].tolist() == [2500, 17, 4000] and ids[0, 3:].tolist() == [i for i in range(K) if i != 17][: K - 3]
    m.close()


@pytest.mark.parametrize("adam_mode", ["dense", "lazy"])
def test_train_tiny_and_repeated_user_batches(pda, c_oracle, adam_mode):
    """B = 1, and a batch where one user repeats (the reference never produces it, a host caller may): the distinct-users
    check must keep such a batch off the fused path and the result must still equal the oracle."""
    rng = np.random.default_rng(2)
    n_users, n_items, d = 50, 40, 16
    U = rng.normal(0, 0.3, (n_users, d)).astype(np.float32)
    I = rng.normal(0, 0.3, (n_items, d)).astype(np.float32)
    m = pda.PDAModel(n_users, n_items, d, train="s_condition", batch_size=8, lr=1e-2, regs=1e-3, init=False)
    m.set_adam_mode(adam_mode)
    m.set_table("user_embedding", U); m.set_table("item_embedding", I)
    ref = c_oracle.CModel(U, I, 1e-2, 1e-3, 8, "s_condition")
    batches = [
        (np.array([3]), np.array([1]), np.array([2]), np.array([0.5], np.float32), np.array([0.25], np.float32)),
        (np.array([4, 9, 4, 4, 7]), np.array([1, 2, 3, 4, 5]), np.array([6, 7, 8, 9, 10]), np.full(5, 0.5, np.float32),
         np.full(5, 0.7, np.float32)),
        (np.array([3, 4]), np.array([11, 12]), np.array([13, 14]), np.ones(2, np.float32), np.ones(2, np.float32)),
    ]
    for b in batches:
        got = m.train_step(*b)
        want = ref.train_step(*[np.asarray(x) for x in b])
        assert np.allclose(got, want, rtol=1e-5, atol=0)
    assert np.abs(m.get_table("user_embedding") - ref.U).max() <= 1e-6 * np.abs(ref.U).max()
    assert np.abs(m.get_table("item_embedding") - ref.I).max() <= 1e-6 * np.abs(ref.I).max()
    m.close()


def test_sampler_with_replacement_and_single_item_users(pda, c_oracle):
    """B > #users with data -> users drawn with replacement (train_new_api.py:387); users with one interaction."""
    from oracle import pda_oracle as po
    n_users, n_items, T, B = 40, 25, 3, 128
    uid = np.arange(0, n_users, 2)
    iid = (uid * 7) % n_items
    t = uid % T
    indptr, items, times = po.build_csr(n_users, uid, iid, t)
    P = po.train_pop_matrix(pop_table(n_items, T, 1), 0.3)
    m = pda.PDAModel(n_users, n_items, 8, train="s_condition", batch_size=B)
    m.set_train_csr(indptr, items, times, unique_times=np.arange(T))
    m.set_train_pop(P)
    got = m.sample_batch(7, 2, 5, B)
    ref = c_oracle.sample_batch(7, 2, 5, B, uid, indptr, items, times, n_items, np.arange(T), P)
    for k in ("users", "pos", "neg", "time"):
        assert np.array_equal(got[k], ref[k]), k
    assert set(got["users"]) <= set(uid.tolist()) and len(set(got["users"])) < B
    assert (got["pos"] == (got["users"] * 7) % n_items).all() and (got["neg"] != got["pos"]).all()
    m.train_sampled(7, 2, 5, 2, B)          # repeated users in the batch: red.global.add on the user rows
    assert np.isfinite(m.read_loss()).all()
    m.close()
