"""Correctness of the GEMM kernel variants (independent CTAs / multicast pair / cta_group::2 pair) against torch on the same
16-bit inputs, then timing on the layer shapes.  COVO_GEMM_MC / COVO_GEMM_CG are read by covo_dbg_gemm on every call."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import covomix_b200  # noqa
from covomix_b200 import _native as nat
dev = torch.device("cuda:0"); P = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
L = nat.lib()
torch.manual_seed(0)

def run(M, N, K, res, bias, outf, outh, act, bn):
    A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    r0 = torch.randn(M, N, device=dev) if res else None
    of = r0.clone() if res else (torch.zeros(M, N, device=dev) if outf else None)
    oh = torch.zeros(M, N, device=dev, dtype=torch.bfloat16) if outh else None
    nat.check(L.covo_dbg_gemm(P(A), P(W), P(b), P(of) if res else None, P(of) if outf else None, P(oh), M, N, K, act, bn, None), "g")
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    if bias: ref = ref + b
    if res: ref = ref + r0
    errs = []
    if outf: errs.append(((of - ref).norm() / ref.norm()).item())
    if outh:
        rh = torch.nn.functional.gelu(ref) if act == 1 else ref
        errs.append(((oh.float() - rh).norm() / rh.norm()).item())
    return max(errs)

bad = 0
for mode in ("base", "cg2"):
    os.environ["COVO_GEMM_CG"] = "2" if mode == "cg2" else "1"
    for (M, N, K, res, bias, outf, outh, act) in [(26400, 1024, 1024, 1, 0, 1, 0, 0), (26400, 3072, 1024, 0, 0, 0, 1, 0),
                                                   (1300, 4096, 1024, 0, 1, 0, 1, 1), (1300, 1024, 4096, 1, 1, 1, 1, 0),
                                                   (256, 1024, 2048, 0, 1, 1, 0, 0), (130, 512, 64, 0, 0, 1, 0, 0),
                                                   (129, 256, 128, 1, 0, 1, 0, 0), (385, 128, 192, 0, 0, 0, 1, 0)]:
        for bn in (256, 128):
            if N % bn: continue
            e = run(M, N, K, res, bias, outf, outh, act, bn)
            ok = e < 6e-3
            bad += not ok
            print(f"{mode:4s} M={M:6d} N={N:5d} K={K:5d} bn={bn} res={res} bias={bias} f32={outf} h={outh} act={act}: rel-L2 {e:.2e} {'ok' if ok else 'FAIL'}", flush=True)
print("FAILED" if bad else "all ok")
if bad: sys.exit(1)
if len(sys.argv) > 1 and sys.argv[1] == "check": sys.exit(0)

M = 26400
for name, N, K, res, bias, outf, outh, act in [("qkv", 3072, 1024, 0, 0, 0, 1, 0), ("out", 1024, 1024, 1, 0, 1, 0, 0),
                                               ("ff1", 4096, 1024, 0, 1, 0, 1, 1), ("ff2", 1024, 4096, 1, 1, 1, 1, 0),
                                               ("skip", 1024, 2048, 0, 1, 1, 0, 0), ("plain-bf16", 4096, 4096, 0, 0, 0, 1, 0)]:
    A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    of = torch.randn(M, N, device=dev) if (outf or res) else None
    oh = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if outh else None
    for mode in ("base", "cg2"):
        os.environ["COVO_GEMM_CG"] = "2" if mode == "cg2" else "1"
        call = lambda: nat.check(L.covo_dbg_gemm(P(A), P(W), P(b), P(of) if res else None, P(of) if outf else None, P(oh), M, N, K, act, 256, None), "g")
        for _ in range(3): call()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:11s} {mode:4s} M={M} N={N} K={K}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
