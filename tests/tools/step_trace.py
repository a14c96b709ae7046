"""Per-phase timeline of the persistent decode-step kernel during one bench-config forward (B=64, V=10000, T=19,
train mode): every CTA stamps %globaltimer at each phase boundary (set_step_trace).  Prints, per phase, the mean
duration over steps 2.. (max over CTAs of the boundary time, successive differences) and the step total."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from show_edit_tell_b200 import _lib, editnet, synth  # noqa: E402
from show_edit_tell_b200.train import XETrainer  # noqa: E402

V, D, A, FD, R, B = 10000, 1024, 512, 2048, 36, int(os.environ.get("DBG_B", "64"))
L = _lib.lib()
torch.manual_seed(0)
dec = editnet.DecoderC(synth.word_map(V), D, D, D, A, FD).cuda()
tr = XETrainer(dec, distributed=False)
b = synth.make_batch(B, V, R, FD, 20, 18, ragged=False, seed=100)
args = [b[k].cuda() for k in ("feats", "caps", "caplens", "prev", "prev_len")]
for _ in range(3):
    tr.step(*args)
torch.cuda.synchronize()
gg, cc = C.c_int(), C.c_int()
L.set_step_geometry(C.byref(gg), C.byref(cc))
G = gg.value
print("persistent grid %d CTAs, clusters of %d" % (G, cc.value))
T = 19
buf = torch.zeros(T * 8 * G + T * 5 * 16 + T * 8, dtype=torch.int64, device="cuda")
L.set_step_trace(C.c_void_p(buf.data_ptr()))
la, st = C.c_longlong(), C.c_longlong()
L.set_step_stats(C.byref(la), C.byref(st), 1)
tr.step(*args)
torch.cuda.synchronize()
L.set_step_trace(None)
L.set_step_stats(C.byref(la), C.byref(st), 1)
print("persistent launches %d covering %d steps" % (la.value, st.value))
tr_ = buf[:T * 8 * G].view(T, 8, G).cpu().double()
fine = buf[T * 8 * G:T * 8 * G + T * 80].view(T, 5, 16).cpu().double()
fenced = buf[T * 8 * G + T * 80:].view(T, 8)[:, :7].cpu().double()
names = ["A lstm", "B h1-consumers", "C1 scores", "C2 context", "D ctx-gate/img", "E copy1", "F copy2"]
end = tr_[:, :7, :].max(dim=2).values          # [T][7] time when the last CTA reached the boundary
first = tr_[:, :7, :].min(dim=2).values
dur = torch.zeros(T, 7)
for t in range(T):
    for p in range(7):
        prev = end[t, p - 1] if p > 0 else (end[t - 1, 6] if t > 0 else float("nan"))
        dur[t, p] = (end[t, p] - prev) / 1e3
print("phase durations (us), mean over steps 2..%d (last-CTA arrival to last-CTA arrival; +barrier latency):" % (T - 1))
for p in range(7):
    print("  %-16s %6.2f   (first-to-last CTA arrival spread %5.2f us)" % (names[p], float(dur[2:, p].mean()),
          float((end[2:, p] - first[2:, p]).mean() / 1e3)))
print("  step total       %6.2f us" % float(dur[2:].sum(1).mean()))

# CTA 0, per GEMM phase: offsets (us) from the completion of the gating barrier (last CTA's arrival stamp)
gate_of = {0: None, 1: 0, 2: 3, 3: 4, 4: 5}
labels = ["gate seen", "first Q", "last Q", "accum ready", "ct0 staged", "staged(bar)", "arrives sent",
          "partners arrived", "finish done"]
cta = int(os.environ.get('SET_STEP_TRACE_CTA', '0'))
own = tr_[:, :7, cta]
print("CTA %d fine timeline, us after the gating barrier completed (mean over steps 2..):" % cta)
for ph, nm in enumerate(["A", "B", "D", "E", "F"]):
    bar_idx = {0: 0, 1: 1, 2: 4, 3: 5, 4: 6}[ph]
    rows = []
    for t in range(2, T):
        g = end[t - 1, 6] if ph == 0 else end[t, gate_of[ph]]
        rows.append([(fine[t, ph, k] - g) / 1e3 for k in (0, 1, 3, 4, 10, 5, 6, 7, 8)] +
                    [(own[t, bar_idx] - g) / 1e3, (fenced[t, bar_idx] - g) / 1e3, (end[t, bar_idx] - g) / 1e3])
    m = torch.tensor(rows).mean(0)
    print("  %s: " % nm + "  ".join("%s %.2f" % (l, float(v)) for l, v in zip(labels + ["own stamp", "proxy-fenced", "last arrival"], m)))

rows = []
for t in range(2, T):
    g = end[t, 1]
    rows.append([(fine[t, 1, k] - g) / 1e3 for k in (11, 12, 13, 14, 15)] + [(own[t, 2] - g) / 1e3, (end[t, 2] - g) / 1e3])
m = torch.tensor(rows).mean(0)
print("  C1 (CTA 0, warp 0): " + "  ".join("%s %.2f" % (l, float(v)) for l, v in zip(
    ["start", "range+prefetch issued", "pass 1 summed", "pass 2 summed", "loop done", "own stamp", "last arrival"], m)))
