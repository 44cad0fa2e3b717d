"""Runs the attention kernel at the C3 shape a few times (target for ncu / timing)."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import covomix_b200  # noqa
from covomix_b200 import _native as nat
dev = torch.device("cuda:0"); P = lambda t: C.c_void_p(t.data_ptr())
Bt, N, H = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 1650, 16
qkv = torch.randn(Bt, N, 3 * H * 64, device=dev).bfloat16(); out = torch.empty(Bt, N, H * 64, device=dev, dtype=torch.bfloat16)
for _ in range(3): nat.check(nat.lib().covo_dbg_attention(P(qkv), P(out), Bt, N, H, 0, None), "a")
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): nat.check(nat.lib().covo_dbg_attention(P(qkv), P(out), Bt, N, H, 0, None), "a")
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
print(f"attention Bt={Bt} N={N} H={H}: {ms*1e3:.1f} us  {4*N*N*64*H*Bt/ms/1e9:.1f} TFLOP/s")
