"""Per-step wall / device durations of the end-to-end loop of bench.py (pinned host inputs through DevicePrefetcher,
loss read back one step behind): shows whether a slow e2e number is spikes or a uniform slowdown."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from show_edit_tell_b200 import editnet, synth
from show_edit_tell_b200.feed import DevicePrefetcher
from show_edit_tell_b200.train import XETrainer
V, D, A, FD, R, B = 10000, 1024, 512, 2048, 36, 64
dev = torch.device("cuda", 0)
torch.manual_seed(0)
dec = editnet.DecoderC(synth.word_map(V), D, D, D, A, FD).to(dev)
tr = XETrainer(dec, distributed=False)
host = synth.make_batch(B, V, R, FD, 20, 18, ragged=False, seed=100, pinned=True)
keys = ("feats", "caps", "caplens", "prev", "prev_len")
stream = torch.cuda.current_stream()
res_host = [torch.zeros(()).pin_memory() for _ in range(2)]
res_ev = [torch.cuda.Event() for _ in range(2)]
mode = os.environ.get("E2E_MODE", "prefetch")
# measure raw H2D bandwidth of the pinned feature tensor
torch.cuda.synchronize()
t0 = time.perf_counter(); x = host["feats"].to(dev, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D of %.1f MB pinned: %.2f ms (%.1f GB/s)" % (host["feats"].numel() * 4 / 1e6, dt * 1e3, host["feats"].numel() * 4 / dt / 1e9))
def loop(n):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    walls = []
    feed = DevicePrefetcher((tuple(host[k] for k in keys) for _ in range(n)), dev)
    evs[0].record(stream)
    for i, batch in enumerate(iter(feed)):
        t0 = time.perf_counter()
        slot = i % 2
        hl = (feed.host_batch[2], feed.host_batch[4]) if mode != "sync_lengths" else None
        loss = tr.step(*batch, host_lengths=hl)
        res_host[slot].copy_(loss.detach(), non_blocking=True)
        res_ev[slot].record(stream)
        evs[i + 1].record(stream)
        if i > 0:
            res_ev[1 - slot].synchronize()
        walls.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    dev_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
    return dev_ms, walls
loop(3)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d, w = loop(20)
    torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / 20 * 1e3
    print("rep %d: %.2f ms/step wall; device gaps (ms): %s" % (rep, tot, " ".join("%.1f" % x for x in d)))
    print("        host ms per iteration: %s" % " ".join("%.1f" % (x * 1e3) for x in w))
