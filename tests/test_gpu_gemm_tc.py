"""tcgen05 3xTF32 GEMM kernel vs fp64 torch, all three operand layouts, skinny ("swap") and square
tiles, K tails, split-K, epilogue options.  fp32-equivalent accuracy is required: max abs error
within 2e-5 of the largest output magnitude -- the level the CUDA-core fp32 FMA chain itself reaches
on these sizes (tools/tc_error_probe.py; the tensor core accumulates round-toward-zero, so its error
is a small systematic shrink instead of a random walk, DESIGN.md "Numerics")."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _L():
    from show_edit_tell_b200 import _lib
    return _lib


def _stats(L, reset=True):
    tc, simt = C.c_longlong(), C.c_longlong()
    L.lib().set_gemm_stats(C.byref(tc), C.byref(simt), int(reset))
    return tc.value, simt.value


SHAPES = [  # (mode, M, N, K)
    (0, 64, 4096, 2048), (0, 64, 4096, 3072), (0, 64, 512, 1024), (0, 37, 1024, 1024), (0, 8, 256, 96),
    (0, 1216, 1000, 1024), (0, 2304, 1024, 2048), (0, 300, 260, 200), (0, 128, 128, 64), (0, 129, 65, 100),
    (1, 64, 4096, 4096), (1, 64, 1024, 512), (1, 1216, 1024, 1000), (1, 40, 2048, 4096),
    (2, 4096, 1024, 1216), (2, 512, 1024, 4608), (2, 1000, 1024, 1216), (2, 96, 160, 70),
]


@pytest.mark.parametrize("mode,M,N,K", SHAPES)
def test_tc_gemm_matches_fp64(mode, M, N, K):
    L = _L()
    lib = L.lib()
    lib.set_gemm_backend(0)
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + mode)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    if mode == 0:      # C[M,N] = A[M,K] W[N,K]^T
        A, Bm, lda, ldb = rnd(M, K), rnd(N, K), K, K
        ref = A.double() @ Bm.double().t()
    elif mode == 1:    # C[M,N] = A[M,K] W[K,N]
        A, Bm, lda, ldb = rnd(M, K), rnd(K, N), K, N
        ref = A.double() @ Bm.double()
    else:              # C[M,N] = A[K,M]^T X[K,N]
        A, Bm, lda, ldb = rnd(K, M), rnd(K, N), M, N
        ref = A.double().t() @ Bm.double()
    bias = rnd(N)
    Cm = rnd(M, N)
    C0 = Cm.clone()
    _stats(L)
    L.check(lib.set_gemm(mode, M, N, K, L.ptr(A), lda, L.ptr(Bm), ldb, L.ptr(bias), L.ptr(Cm), N, 1, 0, None))
    torch.cuda.synchronize()
    tc, simt = _stats(L)
    if mode == 0:
        assert tc == 1 and simt == 0, "tensor-core path was not taken (tc=%d simt=%d)" % (tc, simt)
    else:   # NN / TN operands are MN-major: CUDA-core kernel (the decode path presents them in NT form)
        assert tc == 0 and simt == 1
    ref = ref + bias.double() + C0.double()
    err = float((Cm.double() - ref).abs().max())
    tol = 2e-5 * float(ref.abs().max()) * max(1.0, K / 2048)   # RZ bias grows with the accumulation chain
    print("mode %d %dx%dx%d: max abs err %.3e (tol %.3e)" % (mode, M, N, K, err, tol))
    assert err < tol
    # beta = 0, relu epilogue, vs the CUDA-core kernel bit-for-bit-ish
    C1 = torch.full((M, N), 7.0, device="cuda")
    L.check(lib.set_gemm(mode, M, N, K, L.ptr(A), lda, L.ptr(Bm), ldb, None, L.ptr(C1), N, 0, 1, None))
    ref1 = (ref - bias.double() - C0.double()).clamp_min(0)
    assert float((C1.double() - ref1).abs().max()) < tol


# the big time-batched GEMMs of the benchmarked train step (B=64, T=19, V=10000): more tiles than SMs -> two CTAs per SM
TWIN_SHAPES = [(43776, 512, 1024), (1216, 10000, 1024), (2432, 1024, 2048), (1216, 4096, 1216), (2500, 1000, 300)]


@pytest.mark.parametrize("M,N,K", TWIN_SHAPES)
def test_tc_gemm_twin_configuration_matches_fp64(M, N, K):
    L = _L()
    lib = L.lib()
    lib.set_gemm_backend(0)
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + 7 * K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = A.double() @ W.double().t() + bias.double()
    tol = 2e-5 * float(ref.abs().max()) * max(1.0, K / 2048)
    for rep in range(3):     # repeated: the twin CTAs of an SM interleave differently from launch to launch
        Cm = torch.full((M, N), float("nan"), device="cuda")
        _stats(L)
        lib.set_gemm_twin_launches(1)
        L.check(lib.set_gemm(0, M, N, K, L.ptr(A), K, L.ptr(W), K, L.ptr(bias), L.ptr(Cm), N, 0, 0, None))
        torch.cuda.synchronize()
        tc, simt = _stats(L)
        assert (tc, simt) == (1, 0)
        assert int(lib.set_gemm_twin_launches(1)) == 1, "not launched in the twin configuration"
        err = float((Cm.double() - ref).abs().max())
        assert err < tol, ("twin", M, N, K, rep, err, tol)
    # accumulate (beta = 1) + relu-free epilogue on a second buffer
    C0 = torch.randn(M, N, device="cuda", generator=g)
    C1 = C0.clone()
    L.check(lib.set_gemm(0, M, N, K, L.ptr(A), K, L.ptr(W), K, None, L.ptr(C1), N, 1, 0, None))
    assert float((C1.double() - (ref - bias.double() + C0.double())).abs().max()) < tol


def test_unaligned_problem_falls_back_to_cuda_cores():
    L = _L()
    lib = L.lib()
    lib.set_gemm_backend(0)
    A = torch.randn(64, 130, device="cuda")
    W = torch.randn(53, 130, device="cuda")     # ld = 130 floats -> rows not 16-byte aligned
    Cm = torch.empty(64, 53, device="cuda")
    _stats(L)
    L.check(lib.set_gemm(0, 64, 53, 130, L.ptr(A), 130, L.ptr(W), 130, None, L.ptr(Cm), 53, 0, 0, None))
    tc, simt = _stats(L)
    assert (tc, simt) == (0, 1)
    assert (Cm.double() - A.double() @ W.double().t()).abs().max() < 1e-3
