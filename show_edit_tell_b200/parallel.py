"""Data-parallel plumbing for the train steps (new: the reference is single-process).

One process per GPU, full parameter replicas, the batch sharded by rows.  A step needs exactly one
collective: an all-reduce(SUM) over the flat gradient buffer whose last slot carries the rank's token
count, so gradient sums and the global normaliser travel together (the reference's
loss is a mean over *packed tokens*, editnet.py:575-577, so ranks must contribute sums, not means).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT); returns (rank, world)"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        dist.init_process_group(backend)
    return rank, world


def shard_rows(n_rows, rank, world):
    """contiguous row range of this rank (remainder rows go to the first ranks)"""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sums(flat_grad_with_count, n, local_count, group=None):
    """In place: flat_grad_with_count[:n] holds this rank's gradient of the loss SUM, slot n receives the
    local token count; after the single all-reduce the buffer holds global sums and slot n the global
    count.  Returns the view of the count slot (a device tensor: no host sync)."""
    count_slot = flat_grad_with_count[n:n + 1]
    count_slot.copy_(local_count.reshape(1).to(count_slot.dtype))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad_with_count, op=dist.ReduceOp.SUM, group=group)
    return count_slot
