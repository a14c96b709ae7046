"""Characterise the 3xTF32 kernel's error: growth with K, sign bias (round-toward-zero accumulation
in the tensor core shows up as err anti-correlated with the result's sign)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
torch.manual_seed(0)
M, N = 64, 1024
for dist in ("randn", "uniform", "model-like"):
    for K in (64, 256, 1024, 4096):
        if dist == "randn":
            A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
        elif dist == "uniform":
            A, W = torch.rand(M, K, device="cuda"), torch.rand(N, K, device="cuda")
        else:
            A = torch.tanh(torch.randn(M, K, device="cuda"))
            W = (torch.rand(N, K, device="cuda") * 2 - 1) / 32
        ref = A.double() @ W.double().t()
        res = {}
        for backend in (0, 1):
            lib.set_gemm_backend(backend)
            Cm = torch.empty(M, N, device="cuda")
            L.check(lib.set_gemm(0, M, N, K, L.ptr(A), K, L.ptr(W), K, None, L.ptr(Cm), N, 0, 0, None))
            e = Cm.double() - ref
            res[backend] = (float(e.abs().max()), float((e * ref.sign()).mean()), float(e.abs().mean()))
        print("%-10s K=%5d |C|~%8.2f  tc: max %.2e mean|e| %.2e signed %.2e   simt: max %.2e mean|e| %.2e signed %.2e" % (
            dist, K, float(ref.abs().mean()), res[0][0], res[0][2], res[0][1], res[1][0], res[1][2], res[1][1]))
lib.set_gemm_backend(0)
