import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
def run(mode, M, N, K, beta, act, bias_on, fill=7.0):
    g = torch.Generator(device="cuda").manual_seed(1)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    if mode == 0: A, Bm, lda, ldb = rnd(M, K), rnd(N, K), K, K; ref = A.double() @ Bm.double().t()
    elif mode == 1: A, Bm, lda, ldb = rnd(M, K), rnd(K, N), K, N; ref = A.double() @ Bm.double()
    else: A, Bm, lda, ldb = rnd(K, M), rnd(K, N), M, N; ref = A.double().t() @ Bm.double()
    bias = rnd(N)
    Cm = torch.full((M, N), fill, device="cuda")
    L.check(lib.set_gemm(mode, M, N, K, L.ptr(A), lda, L.ptr(Bm), ldb, L.ptr(bias) if bias_on else None, L.ptr(Cm), N, beta, act, None))
    torch.cuda.synchronize()
    if bias_on: ref = ref + bias.double()
    if act == 1: ref = ref.clamp_min(0)
    if beta: ref = ref + fill
    e = (Cm.double() - ref).abs()
    bad = e > 1e-2
    print("mode %d %dx%dx%d beta %d act %d bias %d: max err %.3e, bad %d/%d" % (mode, M, N, K, beta, act, bias_on, float(e.max()), int(bad.sum()), bad.numel()))
    if bad.any():
        rows = bad.any(1).nonzero().view(-1); cols = bad.any(0).nonzero().view(-1)
        print("   bad rows", rows[:8].tolist(), "...", rows[-3:].tolist(), "n", len(rows), " bad cols", cols[:8].tolist(), "...", cols[-3:].tolist(), "n", len(cols))
        i, j = bad.nonzero()[0].tolist()
        print("   e.g. C[%d,%d] = %.4f ref %.4f   (unrelu'd ref %.4f)" % (i, j, float(Cm[i, j]), float(ref[i, j]), 0.0))
for args in [(0, 64, 4096, 2048, 0, 1, 0), (0, 64, 4096, 2048, 0, 0, 0), (0, 64, 4096, 2048, 1, 1, 0), (0, 64, 512, 1024, 0, 1, 0),
             (1, 64, 1024, 512, 0, 1, 0), (1, 64, 1024, 512, 0, 0, 0), (1, 64, 1024, 512, 1, 0, 1), (1, 1216, 1024, 1000, 0, 1, 0), (1, 1216, 1024, 1000, 0, 0, 0)]:
    run(*args)
