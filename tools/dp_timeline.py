"""timeline of the overlapped gradient all-reduce of the data-parallel train step (bench shape), rank 0:
torchrun --nproc-per-node N tools/dp_timeline.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from show_edit_tell_b200 import editnet, synth
from show_edit_tell_b200.train import XETrainer
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
V, D, A, FD, R, B = 10000, 1024, 512, 2048, 36, 64
torch.manual_seed(0)
dec = editnet.DecoderC(synth.word_map(V), D, D, D, A, FD).to(dev)
tr = XETrainer(dec, distributed=True, trace_overlap=True)
host = synth.make_batch(B, V, R, FD, 20, 18, ragged=False, seed=100 + rank, pinned=False)
batch = tuple(host[k].to(dev) for k in ("feats", "caps", "caplens", "prev", "prev_len"))
hl = (host["caplens"], host["prev_len"])
for i in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record(); tr.step(*batch, host_lengths=hl); e1.record(); torch.cuda.synchronize()
    tl = tr.overlap_timeline()
    if rank == 0 and i >= 4:
        print("step %d: %.2f ms | buckets MB %s | final at %s | all-reduce done at %s | backward ends %.2f" % (
            i, e0.elapsed_time(e1), ["%.0f" % x for x in tl["bucket_mb"]], ["%.2f" % x for x in tl["final_ms"]],
            ["%.2f" % x for x in tl["allreduce_end_ms"]], tl["backward_end_ms"]))
dist.destroy_process_group()
