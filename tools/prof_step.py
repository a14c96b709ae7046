"""One EditNet XE train step (bench.py's workload) inside a cudaProfilerStart/Stop range, for
`ncu --profile-from-start off -k regex:<kernel> -s <n> -c <m> python tools/prof_step.py`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from show_edit_tell_b200 import editnet, synth
from show_edit_tell_b200.train import XETrainer

V, D, A, FD, R, B, CAPW, PREVW = 10000, 1024, 512, 2048, 36, 64, 20, 18
dev = torch.device("cuda", 0)
torch.manual_seed(0)
dec = editnet.DecoderC(synth.word_map(V), D, D, D, A, FD).to(dev)
tr = XETrainer(dec)
b = synth.make_batch(B, V, R, FD, CAPW, PREVW, ragged=False, seed=100)
args = [b[k].to(dev) for k in ("feats", "caps", "caplens", "prev", "prev_len")]
hl = (b["caplens"], b["prev_len"])      # the loader's host copies of the lengths, as bench.py passes them
for _ in range(3):
    tr.step(*args, host_lengths=hl)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(*args, host_lengths=hl)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
