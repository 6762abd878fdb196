"""-m gpu parity: the tcgen05/TMEM GEMM and its fused epilogues vs torch fp32 references of the same ops."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.gpu_util import linear, err_report, dump, device_error


def _mk(M, N, K, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = (torch.randn(N, device="cuda", generator=g) * 0.1).bfloat16()
    return A, W, b


def _check(got, ref, tag, tol=4.5e-3):     # measured on B200: <= 3.4e-3 of the reference maximum (one bf16 rounding of the output)
    rep = err_report(got, ref, tag)
    ok = rep["n_nan"] == 0 and rep["max_abs"] <= tol * max(1.0, rep["ref_absmax"])
    if not ok:
        rep["device_error"] = device_error()
        dump(f"diag_gemm_{tag}.json", rep)
    assert ok, rep


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 128, 64), (128, 256, 128), (256, 256, 256), (300, 192, 96),
                                   (1000, 384, 1536), (8192, 4608, 1536), (666, 1536, 4096), (8192, 64, 1536)])
def test_gemm_bias(M, N, K):
    A, W, b = _mk(M, N, K)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    linear(A, W, b, out)
    torch.cuda.synchronize()
    ref = torch.nn.functional.linear(A.float(), W.float(), b.float())
    _check(out, ref, f"bias_{M}_{N}_{K}")


def test_gemm_identity_layout():
    """A = I: the output must reproduce W^T exactly - catches any swizzle / descriptor / TMEM-lane permutation."""
    K = 64
    A = torch.eye(128, K, device="cuda").bfloat16()
    W = torch.arange(128 * K, device="cuda").reshape(128, K).remainder(251).float().bfloat16()
    out = torch.zeros(128, 128, device="cuda", dtype=torch.bfloat16)
    linear(A, W, None, out)
    torch.cuda.synchronize()
    ref = torch.zeros(128, 128, device="cuda")
    ref[:K] = W.float().t()
    _check(out, ref, "identity", tol=1e-6)


def test_gemm_batched_ragged_rows_and_row_offset():
    """3-D A [batch, rows, K] with rows % 128 != 0 written into a larger joint buffer at a row offset (text rows after image rows)."""
    B, R, K, N, S, off = 3, 77, 128, 192, 200, 100
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(B, R, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    out = torch.zeros(B, S, N, device="cuda", dtype=torch.bfloat16)
    linear(A, W, b, out, num_batch=B, rows_per_batch=R, a_batch_stride=R * K, out_batch_stride=S * N, out_row_offset=off)
    torch.cuda.synchronize()
    ref = torch.zeros(B, S, N, device="cuda")
    ref[:, off:off + R] = torch.nn.functional.linear(A.float(), W.float(), b.float())
    _check(out.reshape(-1, N), ref.reshape(-1, N), "ragged")
    assert float(out[:, :off].abs().max()) == 0 and float(out[:, off + R:].abs().max()) == 0   # nothing outside the range


def test_gemm_gelu():
    A, W, b = _mk(512, 512, 128, seed=4)
    out = torch.empty(512, 512, device="cuda", dtype=torch.bfloat16)
    linear(A, W, b, out, epi=1)
    torch.cuda.synchronize()
    y = torch.nn.functional.linear(A.float(), W.float(), b.float()).bfloat16()
    ref = torch.nn.functional.gelu(y.float(), approximate="tanh")
    _check(out, ref, "gelu")


def test_gemm_gate_residual():
    B, R, K, N = 2, 200, 256, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(B, R, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    gate = torch.randn(B, 3 * N, device="cuda", generator=g).bfloat16()
    h = torch.randn(B, R, N, device="cuda", generator=g).bfloat16()
    h0 = h.clone()
    linear(A, W, b, h, num_batch=B, rows_per_batch=R, a_batch_stride=R * K, out_batch_stride=R * N, epi=2,
           gate=gate[:, N:], gate_batch_stride=3 * N)
    torch.cuda.synchronize()
    y = torch.nn.functional.linear(A.float(), W.float(), b.float()).bfloat16()
    ref = h0 + gate[:, None, N:2 * N] * y
    _check(h.reshape(-1, N), ref.reshape(-1, N), "gate_resid")


def test_gemm_qkv_rmsnorm():
    R, D = 300, 128
    g = torch.Generator(device="cuda").manual_seed(6)
    A = torch.randn(R, D, device="cuda", generator=g).bfloat16()
    W = (torch.randn(3 * D, D, device="cuda", generator=g) / D ** 0.5).bfloat16()
    b = torch.randn(3 * D, device="cuda", generator=g).bfloat16()
    nq = (1 + 0.1 * torch.randn(64, device="cuda", generator=g)).bfloat16()
    nk = (1 + 0.1 * torch.randn(64, device="cuda", generator=g)).bfloat16()
    out = torch.empty(R, 3 * D, device="cuda", dtype=torch.bfloat16)
    linear(A, W, b, out, epi=3, norm_q=nq, norm_k=nk, qk_dim=D)
    torch.cuda.synchronize()
    y = torch.nn.functional.linear(A.float(), W.float(), b.float()).bfloat16()
    q, k, v = y.split(D, dim=1)

    def rms(t, w):
        t = t.reshape(R, -1, 64)
        var = t.float().pow(2).mean(-1, keepdim=True)
        return ((t * torch.rsqrt(var + 1e-6)).bfloat16() * w).reshape(R, -1)
    ref = torch.cat([rms(q, nq), rms(k, nk), v], dim=1)
    _check(out, ref, "qkv_rms")


def test_gemm_rowtable():
    B, R, K, N = 2, 64, 64, 128
    g = torch.Generator(device="cuda").manual_seed(7)
    A = torch.randn(B, R, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    tab = torch.randn(R, N, device="cuda", generator=g)
    out = torch.empty(B, R, N, device="cuda", dtype=torch.bfloat16)
    linear(A, W, b, out, num_batch=B, rows_per_batch=R, a_batch_stride=R * K, out_batch_stride=R * N, epi=4, row_table=tab)
    torch.cuda.synchronize()
    ref = torch.nn.functional.linear(A.float(), W.float(), b.float()).bfloat16().float() + tab
    _check(out.reshape(-1, N), ref.reshape(-1, N), "rowtable")
