"""Runs one GEMM shape a few times (target for ncu).  usage: gemm_one.py N K res bias outf outh act bn"""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import covomix_b200  # noqa
from covomix_b200 import _native as nat
dev = torch.device("cuda:0"); P = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
N, K, res, bias, outf, outh, act, bn = [int(x) for x in sys.argv[1:9]]
M = 26400
A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
b = torch.randn(N, device=dev) if bias else None
of = torch.randn(M, N, device=dev) if (outf or res) else None
oh = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if outh else None
for _ in range(5):
    nat.check(nat.lib().covo_dbg_gemm(P(A), P(W), P(b), P(of) if res else None, P(of) if outf else None, P(oh), M, N, K, act, bn, None), "g")
torch.cuda.synchronize()
