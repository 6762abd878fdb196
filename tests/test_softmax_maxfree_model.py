"""CPU model of the max-free online softmax experiment (flow_factory_b200/csrc/softmax.cuh, -DFFB_ATT_MAXFREE): only the first KV tile
takes its exact row maximum, later tiles keep the reference and move it by a power of two when the running sum passes 2^24.  The model
restates the kernel's state update in fp32 numpy and checks it against an exact softmax on score rows that force several shifts."""
import numpy as np
import pytest


def _online_maxfree(scores, v, sc, tile=64):
    rows, n = scores.shape
    m = np.full(rows, -np.inf, np.float32)
    l = np.zeros(rows, np.float32)
    o = np.zeros((rows, v.shape[1]), np.float32)
    shifts = 0
    for j0 in range(0, n, tile):
        s = scores[:, j0:j0 + tile].astype(np.float32)
        alpha = np.ones(rows, np.float32)
        if np.isinf(m).any():                                   # first tile: exact maximum
            mnew = np.maximum(m, s.max(1))
            alpha = np.exp2((m - mnew) * sc).astype(np.float32)
            m = mnew
        else:
            grow = ~(l <= np.float32(16777216.0))
            if grow.any():
                assert np.isfinite(l).all()
                e = (((l.view(np.uint32) >> 23) & 0xFF).astype(np.int32) - 127)
                e = np.where(grow, e, 0)
                alpha = np.where(grow, ((127 - e).astype(np.uint32) << 23).view(np.float32), np.float32(1.0))
                m = (m + e.astype(np.float32) / np.float32(sc)).astype(np.float32)
                shifts += int(grow.sum())
        p = np.exp2(s * np.float32(sc) - (m * np.float32(sc))[:, None]).astype(np.float32)
        l = (l * alpha + p.sum(1)).astype(np.float32)
        o = o * alpha[:, None] + p @ v[j0:j0 + tile]
    return o / l[:, None], shifts


@pytest.mark.parametrize("trend", [0.0, 0.02, 0.2])
def test_maxfree_matches_exact_softmax(trend):
    rng = np.random.default_rng(0)
    rows, n, d = 16, 64 * 24, 8
    sc = np.float32(0.125 * 1.4426950408889634)
    scores = (rng.standard_normal((rows, n)) * 4 + trend * np.arange(n)[None, :]).astype(np.float32)   # rising scores force reference shifts
    v = rng.standard_normal((n, d)).astype(np.float32)
    got, shifts = _online_maxfree(scores, v, sc)
    z = scores.astype(np.float64) * float(sc)
    p = np.exp2(z - z.max(1, keepdims=True))
    ref = (p / p.sum(1, keepdims=True)) @ v.astype(np.float64)
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-5)
    if trend >= 0.2:
        assert shifts > 0                                        # the shifting path was exercised
