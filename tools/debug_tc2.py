import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
M, N, K = 64, 256, 64
torch.manual_seed(0)
A = torch.randn(M, K, device="cuda"); B = torch.randn(K, N, device="cuda")
ref = A.double() @ B.double()
Cm = torch.full((M, N), 7.0, device="cuda")
L.check(lib.set_gemm(1, M, N, K, L.ptr(A), K, L.ptr(B), N, None, L.ptr(Cm), N, 0, 0, None))
torch.cuda.synchronize()
e = (Cm.double() - ref).abs()
print("xor", os.environ.get("SET_TC_IDESC_XOR"), "max err %.3e" % float(e.max()), "C[0,:4]", Cm[0, :4].tolist(), "ref", ref[0, :4].tolist())
# does C equal some other contraction?
alts = {"A@B": ref, "zeros": torch.zeros_like(ref)}
for k, v in alts.items():
    print("   vs", k, float((Cm.double() - v).abs().max()))
