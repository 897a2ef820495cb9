"""The certificate argument of the tcgen05 filter, replayed in numpy on the oracle's exact scores:
lower bounds of sampled chunks -> tau; upper bounds >= tau -> candidates (+ the 'pop >= tau' branch); a row is certified
when >= K unmasked candidates have exact score >= tau -- and then the exact top-K MUST lie inside the candidate set.
This is a statement about the algorithm (the CUDA kernels are checked against the oracle in tests/test_gpu_eval.py); the
accumulators are emulated as in tests/test_filter_bound.py."""
import numpy as np
import pytest

from test_filter_bound import accumulate, coefficients, split3

TN = 128


def run_filter(U, I, pop, bias, mode, K, se, cw, ordered, masks):
    from oracle import pda_oracle as po
    M, d = U.shape
    N = I.shape[0]
    n_tiles = -(-N // TN)
    cAB, cB = coefficients(d)
    if mode == "condition":
        W, X3, xcol = (I * pop[:, None]).astype(np.float32), split3(pop), pop
    elif bias is not None:
        W, X3, xcol = I, split3(bias), bias
    else:
        W, X3, xcol = I, None, np.zeros(N, np.float32)
    v = accumulate(U, W, X3, chop=False).astype(np.float64)                     # [M, N]
    Y = po.transform_scores(po.exact_scores(U, I), mode, pop if mode == "condition" else None, bias).astype(np.float64)
    un = np.linalg.norm(U.astype(np.float64), axis=1) * 1.00001
    wn = np.linalg.norm(W.astype(np.float64), axis=1) * 1.00001
    pad = n_tiles * TN - N
    tn = np.pad(wn, (0, pad)).reshape(n_tiles, TN).max(axis=1)
    tcol = np.pad(np.abs(xcol.astype(np.float64)), (0, pad)).reshape(n_tiles, TN).max(axis=1)
    E = (cAB * un[:, None] * tn[None, :] + cB * tcol[None, :] + 1e-30) * (1 + 2e-6)      # [M, n_tiles]
    # pass A: the sampled tiles
    n_sel = -(-n_tiles // se)
    if ordered:
        key = tn + tcol
        sel = np.sort(np.argsort(-key, kind="stable")[:n_sel])
    else:
        sel = np.arange(0, n_tiles, se)
    res = []
    for r in range(M):
        keys, dropped = [], 0
        for t in sel:
            for c0 in range(0, TN, cw):
                j = np.arange(t * TN + c0, min(t * TN + c0 + cw, N))
                if len(j) == 0:
                    continue
                if np.isin(j, masks[r]).any():
                    dropped += 1
                    continue
                lb = v[r, j].max() - E[r, t]
                keys.append(lb - abs(lb) * 2e-6)
        if len(keys) < K:
            res.append(None)
            continue
        tau = np.sort(keys)[-K]
        tau = tau - abs(tau) * 1e-5 - 1e-30
        tl = tau - abs(tau) * 2e-6 - 1e-30
        cand = v[r] >= (tl - np.repeat(E[r], TN)[:N])
        if mode == "condition":
            cand |= pop.astype(np.float64) >= tl
        cand[masks[r]] = False
        res.append((tau, np.nonzero(cand)[0], Y[r]))
    return res


@pytest.mark.parametrize("mode,use_bias", [("main_branch", False), ("main_branch", True), ("condition", False)])
@pytest.mark.parametrize("se,cw,ordered", [(1, 32, False), (2, 64, False), (4, 64, True), (8, 64, True)])
@pytest.mark.parametrize("scale", [0.02, 1.0, 6.0])
def test_certified_rows_contain_the_exact_top_k(mode, use_bias, se, cw, ordered, scale):
    rng = np.random.default_rng(int(scale * 100) + se * 7 + len(mode))
    M, N, d, K = 24, 16384 + 77, 64, 20
    U = (rng.normal(0, scale, (M, d)) / np.sqrt(d)).astype(np.float32)
    I = (rng.normal(0, scale, (N, d)) / np.sqrt(d)).astype(np.float32) * (0.3 + rng.random(N) ** 3).astype(np.float32)[:, None]
    pop = (rng.random(N) ** 4).astype(np.float32)
    pop[rng.random(N) < 0.2] = 0.0
    bias = rng.normal(0, 0.2, N).astype(np.float32) if use_bias else None
    S = U.astype(np.float64) @ I.astype(np.float64).T
    # train items: a third of the rows mask their own best items (a fitted model), the rest random ones
    masks = [np.argsort(-S[r])[:60] if r % 3 == 0 else rng.choice(N, rng.integers(0, 40), replace=False) for r in range(M)]
    out = run_filter(U, I, pop, bias, mode, K, se, cw, ordered, masks)
    certified = 0
    for r, o in enumerate(out):
        if o is None:
            continue                                            # fewer than K clean sampled chunks: the exact kernel takes the row
        tau, cand, y = o
        # K clean chunks => K distinct unmasked items with exact score >= tau, all of them candidates: always certified
        assert (y[cand] >= tau).sum() >= K, r
        certified += 1
        yy = y.copy()
        yy[masks[r]] = -np.inf
        top = np.lexsort((np.arange(N), -yy))[:K]               # (score desc, id asc): tf.nn.top_k's order
        assert np.isin(top, cand).all(), (r, np.setdiff1d(top, cand))
    assert certified >= (0.8 * M if se <= 2 else 1), certified
