"""microbenchmark of the GEMM engines on the shapes of the decode path (CUDA events, L2 flushed)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
SHAPES = [("F1 gates1 (swap)", 0, 64, 4096, 2048), ("F5 x2h (swap)", 0, 64, 4096, 3072), ("F2 att (swap)", 0, 64, 512, 1024),
          ("fc fwd", 0, 1216, 10000, 1024), ("features_att fwd", 0, 43776, 512, 1024), ("att_embed", 0, 2304, 1024, 2048),
          ("dX x2h (swap NN)", 1, 64, 4096, 4096), ("dfe_t NN", 1, 43776, 1024, 512), ("dh2raw NN", 1, 1216, 1024, 10000),
          ("dW x2h TN", 2, 4096, 4096, 1216), ("dW feat TN", 2, 512, 1024, 43776), ("dW fc TN", 2, 10000, 1024, 1216)]
for name, mode, M, N, K in SHAPES:
    if mode == 0: A, B, lda, ldb = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), K, K
    elif mode == 1: A, B, lda, ldb = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda"), K, N
    else: A, B, lda, ldb = torch.randn(K, M, device="cuda"), torch.randn(K, N, device="cuda"), M, N
    Cm = torch.zeros(M, N, device="cuda")
    out = []
    for backend in (0, 1):
        lib.set_gemm_backend(backend)
        ts = []
        for it in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.set_gemm(mode, M, N, K, L.ptr(A), lda, L.ptr(B), ldb, None, L.ptr(Cm), N, 0, 0, None))
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out.append(min(ts[1:]))
    fl = 2.0 * M * N * K
    by = 4.0 * (M * K + N * K + M * N)
    print("%-22s mode %d %6dx%6dx%6d  tc %8.1f us (%6.1f TF/s, %6.0f GB/s)   simt %8.1f us (%5.1f TF/s)" % (
        name, mode, M, N, K, out[0] * 1e3, fl / out[0] / 1e9, by / out[0] / 1e6, out[1] * 1e3, fl / out[1] / 1e9))
lib.set_gemm_backend(0)
