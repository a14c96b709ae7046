"""Repeat set_gemm on fixed inputs and look for run-to-run glitches (> tol) per shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
N_IT = int(sys.argv[1]) if len(sys.argv) > 1 else 500
torch.manual_seed(0)
shapes = [(72, 1024, 1024), (72, 4096, 1024), (76, 4096, 1024), (76, 1003, 1024), (4, 4096, 2048), (4, 512, 1024), (4, 4096, 1024),
          (64, 4096, 2048), (64, 1024, 1024), (144, 4096, 1024), (1216, 4096, 1024)]
for (M, N, K) in shapes:
    A = torch.randn(M, K, device="cuda") * 0.1
    W = torch.randn(N, K, device="cuda") * 0.05
    bias = torch.randn(N, device="cuda")
    C = torch.empty(M, N, device="cuda")
    ref = (A.double() @ W.double().t() + bias.double()).float()
    bad = 0; worst = 0.0; first = None
    for it in range(N_IT):
        C.fill_(float("nan"))
        L.check(lib.set_gemm(0, M, N, K, L.ptr(A), K, L.ptr(W), K, L.ptr(bias), L.ptr(C), N, 0, 0, None))
        e = float((C - ref).abs().max()) if not torch.isnan(C).any() else float("inf")
        worst = max(worst, e)
        if e > 1e-4:
            bad += 1
            if first is None:
                d = (C - ref).abs(); d[torch.isnan(d)] = 1e9
                idx = (d > 1e-4).nonzero()
                first = (it, e, idx.shape[0], idx[0].tolist(), idx[-1].tolist())
    print("%5dx%5dx%5d: %d/%d bad, worst err vs fp64 %.3e %s" % (M, N, K, bad, N_IT, worst, first if first else ""))
