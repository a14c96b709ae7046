"""host-side cost per set_gemm call vs GPU time (back-to-back calls, no flush)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
for name, M, N, K in [("tiny", 64, 512, 1024), ("F1", 64, 4096, 2048), ("F5", 64, 4096, 3072), ("fc", 1216, 10000, 1024)]:
    A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    Cm = torch.zeros(M, N, device="cuda")
    for backend in (0, 1):
        lib.set_gemm_backend(backend)
        for _ in range(3):
            L.check(lib.set_gemm(0, M, N, K, L.ptr(A), K, L.ptr(W), K, None, L.ptr(Cm), N, 0, 0, None))
        torch.cuda.synchronize()
        n = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            lib.set_gemm(0, M, N, K, L.ptr(A), K, L.ptr(W), K, None, L.ptr(Cm), N, 0, 0, None)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print("%-5s backend %d: host %.1f us/call, gpu (events, back-to-back, L2-warm) %.1f us/call" % (
            name, backend, (t1 - t0) / n * 1e6, e0.elapsed_time(e1) / n * 1e3))
lib.set_gemm_backend(0)
