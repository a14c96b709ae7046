"""Timeline of the tensor-core launches of one EditNet XE train step (bench.py's workload), taken with
%globaltimer stamps inside the kernels (set_gemm_trace_seq): per launch the grid size, when its first
and last CTA entered, when CTA 0 saw its first operand tile, finished its main loop, finished its
epilogue, and when the last CTA left; and the gap to the next traced launch (= the non-GEMM kernels
in between + launch latency).  Usage: python tools/step_timeline.py [first_launch] [count]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from show_edit_tell_b200 import _lib as L
from show_edit_tell_b200 import editnet, synth
from show_edit_tell_b200.train import XETrainer

V, D, A, FD, R, B, CAPW, PREVW = 10000, 1024, 512, 2048, 36, 64, 20, 18
first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 400
lib = L.lib()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
dec = editnet.DecoderC(synth.word_map(V), D, D, D, A, FD).to(dev)
tr = XETrainer(dec)
b = synth.make_batch(B, V, R, FD, CAPW, PREVW, ragged=False, seed=100)
args = [b[k].to(dev) for k in ("feats", "caps", "caplens", "prev", "prev_len")]
for _ in range(3):
    tr.step(*args)
torch.cuda.synchronize()
STRIDE = 2100 + 8 * 48 + 16
N = first + count
buf = torch.zeros(N * STRIDE, dtype=torch.int64, device=dev)
lib.set_gemm_trace_seq(L.ptr(buf), STRIDE, N)
tr.step(*args)
torch.cuda.synchronize()
lib.set_gemm_trace(None)
t = buf.cpu().view(N, STRIDE)
rows = []
for n in range(N):
    st = t[n, 16:2016:2].double()
    en = t[n, 17:2017:2].double()
    ok = st > 0
    if not bool(ok.any()):
        continue
    rows.append(dict(n=n, ctas=int(ok.sum()), s_first=float(st[ok].min()), s_last=float(st[ok].max()),
                     e_first=float(en[ok].min()), e_last=float(en[ok].max()), stamps=[float(x) for x in t[n, :16]]))
if not rows:
    print("no traced launches")
    sys.exit(0)
t0 = rows[0]["s_first"]
print("launch ctas |  first-in  last-in | cta0: tile-in  loop-done  epi-done | first-out last-out | life  gap-to-next  (us, relative)")
for i, r in enumerate(rows):
    if r["n"] < first:
        continue
    sp = r["stamps"]
    rel = lambda x: (x - r["s_first"]) / 1e3 if x > 0 else float("nan")
    gap = (rows[i + 1]["s_first"] - r["e_last"]) / 1e3 if i + 1 < len(rows) else float("nan")
    nxt_overlap = (rows[i + 1]["s_first"] - r["s_first"]) / 1e3 if i + 1 < len(rows) else float("nan")
    print("%5d %4d | %9.1f %7.1f | %12.1f %9.1f %9.1f | %8.1f %8.1f | %5.1f %6.1f   (next starts +%.1f)" % (
        r["n"], r["ctas"], (r["s_first"] - t0) / 1e3, rel(r["s_last"]), rel(sp[2]), rel(sp[4]), rel(sp[7]),
        rel(r["e_first"]), rel(r["e_last"]), (r["e_last"] - r["s_first"]) / 1e3, gap, nxt_overlap))
    if os.environ.get("EPI"):
        print("        cta0 epilogue: loop-done %.1f accum %.1f staged %.1f partial-written %.1f partners-in %.1f fenced %.1f finished %.1f exit %.1f" % tuple(
            rel(sp[k]) for k in (4, 5, 6, 9, 10, 11, 12, 8)))

# per-K-block SM-clock stamps of CTA 0 for selected launches (KB_STAMP in gemm_tc.cu)
names = ["Q issued", "P issued", "conv saw Q", "conv saw P", "conv done", "mma saw", "mma issued"]
for n in [int(x) for x in os.environ.get("KB_LAUNCHES", "").split(",") if x]:
    kb = t[n, 2100:2100 + 8 * 48].view(48, 8)
    c0 = int(kb[kb > 0].min()) if bool((kb > 0).any()) else 0
    print("launch %d: per-K-block stamps of CTA 0 (SM cycles since the first stamp)" % n)
    print("  kb | " + " | ".join("%10s" % x for x in names))
    for i in range(48):
        if not bool((kb[i] > 0).any()):
            continue
        print("  %2d | " % i + " | ".join("%10d" % (int(kb[i, k]) - c0 if kb[i, k] > 0 else -1) for k in range(7)))
