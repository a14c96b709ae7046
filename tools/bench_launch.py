"""fixed cost of a tensor-core GEMM launch: back-to-back, with/without the split-K memset, and
interleaved with a small-smem kernel (shared-memory carveout switches?)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
M, N, K = 64, 512, 1024
A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
Cm = torch.zeros(M, N, device="cuda")
x = torch.zeros(1024, device="cuda")
def bench(name, fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-50s %.2f us/iter" % (name, e0.elapsed_time(e1) / n * 1e3))
g = lambda beta: lib.set_gemm(0, M, N, K, L.ptr(A), K, L.ptr(W), K, None, L.ptr(Cm), N, beta, 0, None)
bench("tc tiny, beta=1 (no memset)", lambda: g(1))
bench("tc tiny, beta=0 (memset2D + kernel)", lambda: g(0))
bench("torch add_ only", lambda: x.add_(1))
bench("tc tiny beta=1 + torch add_", lambda: (g(1), x.add_(1)))
lib.set_gemm_backend(1)
bench("simt tiny", lambda: g(1))
bench("simt tiny + torch add_", lambda: (g(1), x.add_(1)))
lib.set_gemm_backend(0)
M, N, K = 64, 4096, 2048
A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
Cm = torch.zeros(M, N, device="cuda")
bench("tc F1 64x4096x2048 beta=1", lambda: g(1))
bench("tc F1 beta=1 + torch add_", lambda: (g(1), x.add_(1)))
