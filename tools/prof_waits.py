"""Wait-cycle attribution with the -DFFB_PROFILE build (FFB200_LIB=flow_factory_b200/libffb200_prof.so): where do the TMA /
MMA / epilogue / softmax roles of CTA 0 spend their cycles?  Developer aid, not a bench number."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("FFB200_LIB", os.path.join(ROOT, "flow_factory_b200", "libffb200_prof.so"))
sys.path.insert(0, ROOT)
import torch
from flow_factory_b200 import _lib, ops

NAMES = {0x10: "gemm.tma wait empty", 0x20: "gemm.mma wait tmem_empty", 0x21: "gemm.mma wait full", 0x30: "gemm.epi wait tmem_full (4 warps)",
         0x70: "gemm.tma total", 0x71: "gemm.mma total", 0x72: "gemm.epi0 total", 0x73: "gemm.epi1 total", 0x74: "gemm.epi2 total", 0x75: "gemm.epi3 total",
         0x40: "attn.tma wait k_empty", 0x41: "attn.tma wait v_empty", 0x50: "attn.mma wait k_full", 0x51: "attn.mma wait s_empty",
         0x52: "attn.mma wait q_full", 0x53: "attn.mma wait v_full", 0x54: "attn.mma wait p_full", 0x55: "attn.mma wait o_empty",
         0x60: "attn.softmax wait s_full (8 warps)", 0x61: "attn.softmax wait o_full (8 warps)",
         0x62: "attn.softmax phase: ld S (8 warps)", 0x63: "attn.softmax phase: max", 0x64: "attn.softmax phase: exp+sum+pack",
         0x65: "attn.softmax phase: O readback+accumulate", 0x66: "attn.softmax phase: P store+fence+arrive", 0x67: "attn.softmax phase: loop", 0x68: "attn.softmax phase: wait s_full",
         0x78: "attn.softmax w0 total", 0x79: "attn.softmax w1 total", 0x7c: "attn.softmax w4 total"}


def read():
    buf = (C.c_ulonglong * 256)()
    L = _lib.lib()
    L.ffb200_debug_read_prof.argtypes = [C.POINTER(C.c_ulonglong * 256), C.c_int]
    L.ffb200_debug_read_prof(C.byref(buf), 256)
    return {NAMES.get(i, hex(i)): [int(buf[i]), int(buf[128 + i])] for i in range(128) if buf[i]}


def main():
    torch.manual_seed(0)
    res = {}
    for name, (M, N, K, epi) in {"qkv": (8192, 4608, 1536, 3), "mlp_up": (8192, 6144, 1536, 1), "attn_out": (8192, 1536, 1536, 2)}.items():
        A = torch.randn(M, K, device="cuda").bfloat16(); W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16(); o = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        kw = dict(epi=epi)
        if epi == 2: kw.update(gate=torch.randn(1, N, device="cuda").bfloat16(), gate_batch_stride=N)
        if epi == 3: kw.update(norm_q=torch.ones(64, device="cuda").bfloat16(), norm_k=torch.ones(64, device="cuda").bfloat16(), qk_dim=N // 3)
        for _ in range(3): ops.linear(A, W, b, o, **kw)
        torch.cuda.synchronize(); read()
        ops.linear(A, W, b, o, **kw); torch.cuda.synchronize()
        res["gemm_" + name] = read()
    B, S, H = 2, 4429, 24
    qkv = torch.randn(B, S, 3 * 64 * H, device="cuda").bfloat16()
    for _ in range(3): ops.attention(qkv, H)
    torch.cuda.synchronize(); read()
    ops.attention(qkv, H); torch.cuda.synchronize()
    res["attention"] = read()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
