import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
buf = torch.zeros(16 + 2048, dtype=torch.int64, device="cuda")
lib.set_gemm_trace(L.ptr(buf))
x = torch.zeros(1 << 20, device="cuda")
for (M, N, K) in [(64, 4096, 2048), (64, 4096, 3072)]:
    A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    Cm = torch.zeros(M, N, device="cuda")
    for it in range(3):
        buf.zero_()
        x.add_(1.0)          # a small kernel right before, like the pointwise kernels of the chain
        L.check(lib.set_gemm(0, M, N, K, L.ptr(A), K, L.ptr(W), K, None, L.ptr(Cm), N, 1, 0, None))
        torch.cuda.synchronize()
    t = buf.cpu()
    starts = t[16::2][:128].double(); ends = t[17::2][:128].double()
    ok = starts > 0
    s0 = starts[ok].min()
    print("%dx%dx%d: CTAs %d; start spread %.2f us (max-min), lifetime min/mean/max %.2f/%.2f/%.2f us, first start -> last end %.2f us" % (
        M, N, K, int(ok.sum()), float(starts[ok].max() - s0) / 1e3, float((ends - starts)[ok].min()) / 1e3,
        float((ends - starts)[ok].mean()) / 1e3, float((ends - starts)[ok].max()) / 1e3, float(ends[ok].max() - s0) / 1e3))
