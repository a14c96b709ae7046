"""all-reduce time of the trainer's gradient buckets on this node, alone (no compute underneath): what the data-parallel
step has to hide.  torchrun --nproc-per-node N tools/nccl_allreduce_probe.py"""
import os, sys
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sizes_mb = [41, 79, 101, 78, 39, 338]
bufs = [torch.zeros(int(mb * 1e6 / 4), device="cuda") for mb in sizes_mb]
for b in bufs: dist.all_reduce(b)
torch.cuda.synchronize()
for mb, b in zip(sizes_mb, bufs):
    ts = []
    for it in range(5):
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dist.all_reduce(b); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    if rank == 0:
        print("all-reduce %4d MB on %d GPUs: %.3f ms  (algbw %.0f GB/s, busbw %.0f GB/s)" % (mb, world, t, mb / t, mb / t * 2 * (world - 1) / world))
dist.destroy_process_group()
