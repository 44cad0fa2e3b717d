"""Per-tile cost of the narrow implicit-GEMM shapes of the vocoder (plain K instead of taps: same loads, same epilogue):
M rows x N channels x K = taps * C_in, 16-bit output only (conv1 of a ResBlock) and fp32 + residual + 16-bit (conv2)."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import covomix_b200  # noqa
from covomix_b200 import _native as nat
dev = torch.device("cuda:0"); P = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
L = nat.lib()
def run(M, N, K, res, bn):
    A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
    of = torch.randn(M, N, device=dev) if res else None
    oh = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    call = lambda: nat.check(L.covo_dbg_gemm(P(A), P(W), None, P(of) if res else None, P(of) if res else None, P(oh), M, N, K, 2, bn, None), "g")
    for _ in range(3): call()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): call()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
    tiles = (M + 127) // 128 * (N // bn)
    per_tile_us = ms * 1e3 * 148 / tiles
    byts = M * K * 2 + M * N * 2 + (M * N * 8 if res else 0)
    print(f"M={M} N={N} K={K:4d} bn={bn:3d} {'f32+res+h' if res else 'h only  '}: {ms*1e3:7.1f} us  {2*M*N*K/ms/1e9:6.1f} TFLOP/s  "
          f"{per_tile_us:5.2f} us/tile/SM  min DRAM {byts/ms/1e6:6.0f} GB/s", flush=True)
for (M, N, bn) in ((960000, 64, 64), (240000, 128, 128), (60000, 256, 256), (60000, 256, 128)):
    for K in (N, 3 * N, 7 * N, 11 * N):
        for res in (0, 1):
            run(M, N, K, res, bn)
