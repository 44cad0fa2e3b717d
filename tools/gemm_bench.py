"""Micro-benchmark of the tcgen05 GEMM kernel on the velocity net's layer shapes (C3: M = 26400)."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import covomix_b200  # noqa
from covomix_b200 import _native as nat
dev = torch.device("cuda:0"); P = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
L = nat.lib()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 26400
for name, N, K, res, bias, outf, outh, act in [("qkv", 3072, 1024, 0, 0, 0, 1, 0), ("out", 1024, 1024, 1, 0, 1, 0, 0),
                                               ("ff1", 4096, 1024, 0, 1, 0, 1, 1), ("ff2", 1024, 4096, 1, 1, 1, 1, 0),
                                               ("skip", 1024, 2048, 0, 1, 1, 0, 0), ("plain-f32", 4096, 4096, 0, 0, 1, 0, 0),
                                               ("plain-bf16", 4096, 4096, 0, 0, 0, 1, 0)]:
    A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    of = torch.randn(M, N, device=dev) if (outf or res) else None
    oh = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if outh else None
    for bn in (256, 128):
        call = lambda: nat.check(L.covo_dbg_gemm(P(A), P(W), P(b), P(of) if res else None, P(of) if outf else None, P(oh), M, N, K, act, bn, None), "g")
        for _ in range(3): call()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:11s} M={M} N={N} K={K} bn={bn}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
    t = torch.matmul  # cuBLAS reference point
    for _ in range(3): t(A, W.t())
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): t(A, W.t())
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
    print(f"{'  cublas':11s} {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
