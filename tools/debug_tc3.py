import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
def run(mode, M, N, K, lda=None, ldb=None, beta=1):
    g = torch.Generator(device="cuda").manual_seed(1)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    if mode == 2:
        lda = lda or M; ldb = ldb or N
        A = rnd(K, lda); Bm = rnd(K, ldb)
        ref = A[:, :M].double().t() @ Bm[:, :N].double()
    Cm = torch.full((M, N), 7.0, device="cuda")
    import ctypes as C
    tc, simt = C.c_longlong(), C.c_longlong()
    lib.set_gemm_stats(C.byref(tc), C.byref(simt), 1)
    L.check(lib.set_gemm(mode, M, N, K, L.ptr(A), lda, L.ptr(Bm), ldb, None, L.ptr(Cm), N, beta, 0, None))
    torch.cuda.synchronize()
    lib.set_gemm_stats(C.byref(tc), C.byref(simt), 1)
    if beta: ref = ref + 7.0
    e = (Cm.double() - ref).abs()
    print("mode %d %dx%dx%d lda %d ldb %d: tc=%d max err %.3e  C[0,:3]=%s ref=%s" % (mode, M, N, K, lda, ldb, tc.value, float(e.max()), Cm[0,:3].tolist(), ref[0,:3].tolist()))
for a in [(2, 32, 32, 108), (2, 32, 32, 128), (2, 64, 64, 128), (2, 128, 64, 128), (2, 128, 128, 128), (2, 4096, 1024, 144), (2, 4096, 1024, 1216), (2, 1024, 1024, 152, 2560, 1024),
          (2, 512, 1024, 152, 2560, 4096), (2, 1024, 1024, 1216, 1024, 4096)]:
    run(*a)
