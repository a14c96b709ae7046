"""per-K-block stamps (clock64 of CTA 0) of one tensor-core GEMM launch: where a K-block's time goes
slots: 0 Q load issued, 1 P load issued, 2 Q landed (converter saw it), 3 P landed, 4 conversion done, 5 MMA issue, 6 MMA issued"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from show_edit_tell_b200 import _lib as L
lib = L.lib()
shapes = [(64, 1024, 1024, 1), (64, 3072, 4096, 0), (64, 1024, 4096, 1), (64, 1024, 8192, 1)]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    # the time-batched shapes; run with SET_TC_BIG=0 to trace the one-tile-per-CTA kernels (the persistent kernel keeps
    # no per-K-block stamps)
    shapes = [(43776, 512, 1024, 0), (1216, 10000, 1024, 0), (2304, 1024, 2048, 0)]
for M, N, K, beta in shapes:
    A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    Cm = torch.zeros(M, N, device="cuda")
    buf = torch.zeros(16 + 4096, dtype=torch.int64, device="cuda")
    for it in range(3):
        buf.zero_()
        lib.set_gemm_trace(L.ptr(buf))
        L.check(lib.set_gemm(0, M, N, K, L.ptr(A), K, L.ptr(W), K, None, L.ptr(Cm), N, beta, 0, None))
        torch.cuda.synchronize()
    lib.set_gemm_trace(None)
    ph = buf[:13].cpu().double()
    st = buf[16:16 + 2000:2].cpu().double(); en = buf[17:17 + 2000:2].cpu().double()
    live = st > 0
    g0 = float(st[live].min())
    print("shape %dx%dx%d beta %d phase stamps of CTA 0 (us after the first CTA start): %s" % (M, N, K, beta, " ".join(
        "%d:%.1f" % (i, (float(ph[i]) - g0) / 1e3) for i in range(13) if ph[i] > 0)))
    print("   CTAs traced %d: start min/median/max %.1f/%.1f/%.1f us, end min/median/max %.1f/%.1f/%.1f us" % (
        int(live.sum()), 0.0, (float(st[live].median()) - g0) / 1e3, (float(st[live].max()) - g0) / 1e3,
        (float(en[live].min()) - g0) / 1e3, (float(en[live].median()) - g0) / 1e3, (float(en[live].max()) - g0) / 1e3))
    t = buf[2100:2100 + 8 * 48].view(48, 8).cpu().double()
    nkb = min(48, K // 32)
    t0 = float(t[0, 1]) if t[0, 1] > 0 else float(t[0, 0])
    print("shape %dx%dx%d  (cycles relative to the first load; CTA 0)" % (M, N, K))
    print("  kb   Qissue  Pissue  Qland   Pland  convdone MMAiss  | MMA-to-MMA")
    prev = None
    for i in range(nkb):
        r = [(float(t[i, s]) - t0) if t[i, s] > 0 else float('nan') for s in range(7)]
        gap = (r[5] - prev) if prev is not None else float('nan')
        prev = r[5]
        print("  %2d  %7.0f %7.0f %7.0f %7.0f %7.0f %7.0f  | %6.0f" % (i, r[0], r[1], r[2], r[3], r[4], r[5], gap))
