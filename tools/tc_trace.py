import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
lib.set_gemm_trace(L.ptr(buf))
names = ["start", "setup done", "first TMA landed", "first stage converted", "last stage converted", "accumulator ready", "staged to smem", "epilogue done", "after final sync"]
for (M, N, K) in [(64, 512, 1024), (64, 4096, 2048), (1216, 1024, 1024)]:
    A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    Cm = torch.zeros(M, N, device="cuda")
    for it in range(3):
        buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.set_gemm(0, M, N, K, L.ptr(A), K, L.ptr(W), K, None, L.ptr(Cm), N, 1, 0, None))
        e1.record(); torch.cuda.synchronize()
    t = buf.cpu().tolist()
    print("%dx%dx%d  event time %.1f us" % (M, N, K, e0.elapsed_time(e1) * 1e3))
    for i, n in enumerate(names):
        print("    %-24s +%7.2f us" % (n, (t[i] - t[0]) / 1e3))
