"""clock64 timeline of CTA 0 for the K = 1024 layer GEMMs with progressively richer epilogues (COVO_GEMM_TRACE=1)."""
import ctypes as C, os, sys, torch
os.environ["COVO_GEMM_TRACE"] = "1"
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
    print(name, file=sys.stderr, flush=True)
    for _ in range(2):
        nat.check(L.covo_dbg_gemm(P(A), P(W), P(b), P(of) if res else None, P(of) if outf else None, P(oh), M, N, K, act, 256, None), "g")
run("bf16 out N=4096 K=1024", 4096, 1024, 0, 0, 0, 1, 0)
run("bf16 out + bias", 4096, 1024, 0, 1, 0, 1, 0)
run("bf16 out + bias + gelu (ff1)", 4096, 1024, 0, 1, 0, 1, 1)
run("bf16 out N=4096 K=4096", 4096, 4096, 0, 0, 0, 1, 0)
run("f32 out + residual (out-proj) N=1024 K=1024", 1024, 1024, 1, 0, 1, 0, 0)
run("f32 + residual + bias + bf16 (ff2) N=1024 K=4096", 1024, 4096, 1, 1, 1, 1, 0)
