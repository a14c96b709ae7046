"""Host-side cost of enqueuing one train step (how far the host runs ahead of the GPU)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from show_edit_tell_b200 import editnet, synth
from show_edit_tell_b200.train import XETrainer
V, D, A, FD, R, B = 10000, 1024, 512, 2048, 36, 64
torch.manual_seed(0)
dec = editnet.DecoderC(synth.word_map(V), D, D, D, A, FD).cuda()
tr = XETrainer(dec, distributed=False)
b = synth.make_batch(B, V, R, FD, 20, 18, ragged=False, seed=100)
args = [b[k].cuda() for k in ("feats", "caps", "caplens", "prev", "prev_len")]
for _ in range(3):
    tr.step(*args)
torch.cuda.synchronize()
ts = []
for _ in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); tr.step(*args); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t0))
print("host enqueue ms / total ms per step:", ["%.2f/%.2f" % (a * 1e3, c * 1e3) for a, c in ts])
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    tr.step(*args)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
