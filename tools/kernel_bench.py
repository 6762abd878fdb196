"""Micro-benchmarks of the two tensor-core kernels at SD3.5-medium 1024^2 shapes (CUDA events, L2 flushed between reps).
Prints one JSON line per case; cuBLAS / torch SDPA are timed beside them as the library reference."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.gpu_util import linear, ptr, stream
from flow_factory_b200 import _lib

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    out = []
    for (M, N, K, epi, name) in [(8192, 4608, 1536, 3, "qkv"), (8192, 1536, 1536, 2, "attn_out"), (8192, 6144, 1536, 1, "mlp_up"),
                                 (8192, 1536, 6144, 2, "mlp_down"), (666, 1536, 4096, 0, "ctx_embed"), (32768, 4608, 1536, 3, "qkv_b8"),
                                 (65536, 4608, 1536, 3, "qkv_b16"), (65536, 1536, 1536, 2, "attn_out_b16"), (65536, 1536, 6144, 2, "mlp_down_b16"), (65536, 6144, 1536, 1, "mlp_up_b16")]:
        A = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        o = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        gate = torch.randn(1, N, device="cuda").bfloat16()
        nq = torch.ones(64, device="cuda").bfloat16()
        kw = dict(epi=epi)
        if epi == 2: kw.update(gate=gate, gate_batch_stride=N)
        if epi == 3: kw.update(norm_q=nq, norm_k=nq, qk_dim=N // 3)
        ms = timeit(lambda: linear(A, W, b, o, **kw))
        ms_lib = timeit(lambda: torch.nn.functional.linear(A, W, b))
        fl = 2.0 * M * N * K
        out.append(dict(kernel="gemm", name=name, M=M, N=N, K=K, ms=ms, tflops=fl / ms / 1e9, cublas_ms=ms_lib, cublas_tflops=fl / ms_lib / 1e9))
        print(json.dumps(out[-1]), flush=True)
    for (B, S, H) in [(2, 4429, 24), (2, 4096, 24), (8, 4429, 24)]:
        qkv = torch.randn(B, S, 3 * 64 * H, device="cuda").bfloat16()
        o = torch.empty(B, S, 64 * H, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: _lib.check(_lib.lib().ffb200_attention(ptr(qkv), B, S, H, ptr(o), stream())))   # general path: unscaled keys
        from flow_factory_b200 import ops
        qkv_pre = qkv.float()
        qkv_pre[..., 64 * H: 128 * H] *= 64 ** -0.5 * 1.4426950408889634      # the engines' layout: keys pre-scaled in the QKV GEMM epilogue
        qkv_pre = qkv_pre.bfloat16()
        ms_pre = timeit(lambda: ops.attention(qkv_pre, H, o, k_prescaled=True))
        q, k, v = [t.reshape(B, S, H, 64).transpose(1, 2) for t in qkv.split(64 * H, dim=2)]
        ms_lib = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
        fl = 4.0 * B * H * S * S * 64
        out.append(dict(kernel="attention", B=B, S=S, H=H, ms=ms, tflops=fl / ms / 1e9, engine_layout_ms=ms_pre, engine_layout_tflops=fl / ms_pre / 1e9,
                        sdpa_ms=ms_lib, sdpa_tflops=fl / ms_lib / 1e9))
        print(json.dumps(out[-1]), flush=True)


def flux_attention():
    """head_dim 128 at the FLUX.1 1024^2 shape: S = 512 text + 4096 image tokens, 24 heads."""
    from flow_factory_b200 import ops
    for (B, S, H) in [(1, 4608, 24), (4, 4608, 24)]:
        qkv = torch.randn(B, S, 3 * 128 * H, device="cuda").bfloat16()
        o = torch.empty(B, S, 128 * H, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.attention(qkv, H, o, head_dim=128))                  # general path: unscaled keys
        qkv_pre = qkv.float()
        qkv_pre[..., 128 * H: 256 * H] *= 128 ** -0.5 * 1.4426950408889634           # the engines' layout (keys pre-scaled in the QKV GEMM epilogue)
        qkv_pre = qkv_pre.bfloat16()
        ms_pre = timeit(lambda: ops.attention(qkv_pre, H, o, head_dim=128, k_prescaled=True))
        q, k, v = [t.reshape(B, S, H, 128).transpose(1, 2) for t in qkv.split(128 * H, dim=2)]
        ms_lib = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
        fl = 4.0 * B * H * S * S * 128
        print(json.dumps(dict(kernel="attention_d128", B=B, S=S, H=H, ms=ms, tflops=fl / ms / 1e9, engine_layout_ms=ms_pre,
                              engine_layout_tflops=fl / ms_pre / 1e9, sdpa_ms=ms_lib, sdpa_tflops=fl / ms_lib / 1e9)), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "flux":
        flux_attention(); sys.exit(0)
    main()
