"""GPU-side gap between consecutive tensor-core GEMM launches (and with a small kernel in between)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
M, N, K = 64, 4096, 2048
A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
Cm = torch.zeros(M, N, device="cuda")
x = torch.zeros(1 << 16, device="cuda")
n = 24
bufs = [torch.zeros(16 + 2048, dtype=torch.int64, device="cuda") for _ in range(n)]
for mode in ("tc back-to-back", "tc + small torch kernel between", "tc + cells kernel (dropout_keep_mask) between"):
    for b in bufs: b.zero_()
    tmp = torch.empty(1 << 16, device="cuda")
    torch.cuda.synchronize()
    for i in range(n):
        lib.set_gemm_trace(L.ptr(bufs[i]))
        lib.set_gemm(0, M, N, K, L.ptr(A), K, L.ptr(W), K, None, L.ptr(Cm), N, 1, 0, None)
        if mode.startswith("tc + small"): x.add_(1.0)
        if mode.startswith("tc + cells"): lib.set_dropout_keep_mask(L.ptr(tmp), 1 << 16, 1, 1, 0, None)
    torch.cuda.synchronize()
    st = [float(b[16::2][:128].double().min()) for b in bufs]
    en = [float(b[17::2][:128].double().max()) for b in bufs]
    gaps = [(st[i + 1] - en[i]) / 1e3 for i in range(8, n - 1)]
    life = [(en[i] - st[i]) / 1e3 for i in range(8, n)]
    print("%-48s grid lifetime %.1f us, gap to next TC grid: median %.1f us (min %.1f max %.1f)" % (
        mode, sorted(life)[len(life) // 2], sorted(gaps)[len(gaps) // 2], min(gaps), max(gaps)))
