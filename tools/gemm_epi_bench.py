"""Epilogue cost on the K=1024 layer shapes: same GEMM with progressively richer epilogues (M = 26400)."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import covomix_b200  # noqa
from covomix_b200 import _native as nat
dev = torch.device("cuda:0"); P = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
L = nat.lib(); M = 26400
def run(name, N, K, res, bias, outf, outh, act):
    A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    of = torch.randn(M, N, device=dev) if (outf or res) else None
    oh = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if outh else None
    call = lambda: nat.check(L.covo_dbg_gemm(P(A), P(W), P(b), P(of) if res else None, P(of) if outf else None, P(oh), M, N, K, act, 256, None), "g")
    for _ in range(3): call()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): call()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 20
    print(f"{name:34s} N={N} K={K}: {ms*1e3:7.1f} us {2*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
for K in (1024, 2048, 4096):
    run("bf16 out", 4096, K, 0, 0, 0, 1, 0)
run("bf16 out + bias", 4096, 1024, 0, 1, 0, 1, 0)
run("bf16 out + bias + gelu (ff1)", 4096, 1024, 0, 1, 0, 1, 1)
run("f32 out", 1024, 1024, 0, 0, 1, 0, 0)
run("f32 out + residual (out-proj)", 1024, 1024, 1, 0, 1, 0, 0)
run("f32 out", 1024, 4096, 0, 0, 1, 0, 0)
run("f32 + residual + bias + bf16 (ff2)", 1024, 4096, 1, 1, 1, 1, 0)
run("bf16 out N=3072 (qkv w/o rope)", 3072, 1024, 0, 0, 0, 1, 0)
A = torch.randn(M, 1024, device=dev).bfloat16(); W = torch.randn(4096, 1024, device=dev).bfloat16()
for _ in range(3): torch.matmul(A, W.t())
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
for _ in range(20): torch.matmul(A, W.t())
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 20
print(f"{'cuBLAS bf16 N=4096 K=1024':34s}: {ms*1e3:7.1f} us {2*M*4096*1024/ms/1e9:7.1f} TFLOP/s")
