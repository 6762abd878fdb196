"""CPU models of the attention softmax (flow_factory_b200/csrc/softmax.cuh, attention.cu), fp32 numpy / torch in the device code's operation order:
  * the max-free online softmax: only the first KV tile takes its exact row maximum, later tiles keep the reference and move it by an exact
    power of two when the running sum passes a threshold (the model shifts at 2^24 so that a short row exercises the path; the kernels at
    2^64) - checked against an exact softmax on score rows that force several shifts;
  * the polynomial exp2 of the FMA-pipe slots: accuracy inside its valid range, garbage outside it;
  * the range proof that lets the head_dim-64 kernel drop the per-tile range check: RMS-normed heads bound every score."""
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


# ---------------------------------------------------------------------------------------------------------------------------------------
# CPU restatements of two more claims of softmax.cuh / attention.cu (fp32 numpy, same operation order as the device code)
# ---------------------------------------------------------------------------------------------------------------------------------------
def _exp2_poly(x):
    """exp2_poly_pair of softmax.cuh: Cody-Waite split through the 1.5 * 2^23 magic add, degree-3 minimax polynomial, exponent insert."""
    x = x.astype(np.float32)
    magic = np.float32(12582912.0)
    xr = (x + magic).astype(np.float32)
    r = (xr - magic).astype(np.float32)
    f = (x - r).astype(np.float32)
    p = (f * np.float32(0.05517132207751274) + np.float32(0.24261054396629333)).astype(np.float32)
    p = (p * f + np.float32(0.6932609677314758)).astype(np.float32)
    p = (p * f + np.float32(0.9999281167984009)).astype(np.float32)
    bits = (p.view(np.uint32).astype(np.uint64) + ((xr.view(np.uint32).astype(np.uint64) << 23) & 0xFFFFFFFF)) & 0xFFFFFFFF
    return bits.astype(np.uint32).view(np.float32)


def test_polynomial_exp2_accuracy_and_valid_range():
    """Max relative error 7.5e-5 inside +-126 (far below the 2^-9 of P's bf16 rounding); beyond that the exponent insert wraps - which
    is why the kernel either checks the range per tile or proves it from the RMSNorm weights."""
    x = np.linspace(-125.5, 126.0, 2_000_001).astype(np.float32)
    rel = np.abs(_exp2_poly(x).astype(np.float64) / np.exp2(x.astype(np.float64)) - 1.0)
    assert rel.max() <= 7.6e-5
    edge = _exp2_poly(np.array([-126.0], np.float32)).astype(np.float64) / np.exp2(-126.0)      # the last binade before the wrap: the result is
    assert abs(edge[0] - 1.0) <= 2e-4                                                              # subnormal (1e-38: flushed by P's bf16 anyway)
    bad = _exp2_poly(np.array([130.0, -130.0], np.float32))
    ref = np.exp2(np.array([130.0, -130.0], np.float64))
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        assert not np.allclose(bad.astype(np.float64), ref, rtol=1e-2)          # garbage outside the range, as documented


@pytest.mark.parametrize("w_scale", [0.5, 1.0, 2.5])
def test_rmsnorm_range_proof_bounds_every_score(w_scale):
    """attention.cu prologue: q = bf16(bf16(x rsqrt(mean x^2 + eps)) w_q), k' = bf16(bf16(bf16(...) w_k) scale log2e)  =>
    |q . k'| <= 64 * 1.016 * max|w_q| max|w_k| scale log2e (Cauchy-Schwarz), for adversarially aligned heads too."""
    import torch
    g = torch.Generator().manual_seed(5)
    d, n = 64, 4096
    sl2 = d ** -0.5 * 1.4426950408889634
    wq = (w_scale * (0.5 + 0.5 * torch.rand(d, generator=g))).bfloat16()
    wk = (w_scale * (0.5 + 0.5 * torch.rand(d, generator=g))).bfloat16()

    def rms(x, w):
        n_ = (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16().float()
        return (n_ * w.float()).bfloat16().float()
    x = torch.randn(n, d, generator=g) * torch.logspace(-3, 3, n)[:, None]              # any input scale: RMSNorm removes it
    q = rms(x, wq)
    aligned = torch.sign(wq.float() * wk.float())[None, :] * x                            # keys aligned with their query: the worst case
    k = (rms(torch.cat([aligned, torch.randn(n, d, generator=g)]), wk) * sl2).bfloat16().float()
    scores = q @ k.t()
    bound = 64 * 1.016 * float(wq.float().abs().max()) * float(wk.float().abs().max()) * sl2
    assert float(scores.abs().max()) <= bound
    assert float(scores.abs().max()) >= 0.3 * bound * (0.5 ** 2)                          # and the bound is not vacuous for aligned heads
