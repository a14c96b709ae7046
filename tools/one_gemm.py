import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
M, N, K = [int(x) for x in sys.argv[1:4]]
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
Cm = torch.zeros(M, N, device="cuda")
for _ in range(3):
    L.check(lib.set_gemm(mode, M, N, K, L.ptr(A), K, L.ptr(W), K, None, L.ptr(Cm), N, 0, 0, None))
torch.cuda.synchronize()
