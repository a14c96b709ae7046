"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2: sharding the batch by
rows, contributing loss SUMS plus a token count through ONE all-reduce, and normalising by the global
count reproduces the single-process gradient of the mean loss on the concatenated batch (SURVEY.md
§8e).  The per-shard compute here is the CPU oracle (test infrastructure); the CUDA kernels are
exercised by the -m gpu tests and by `bench.py --gpus N`."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import editnet_oracle as EO
from oracle import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


CFG = dict(V=41, D=16, A=8, Fdim=32, R=5, cap_width=8, prev_width=6, B=6)


def _shard_sum_grads(sd, batch, lo, hi):
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    sub = {k: v[lo:hi] for k, v in batch.items()}
    preds, caps_sorted, dl, _ = EO.xe_forward(sd, sub["feats"], sub["caps"], sub["caplens"], sub["prev"], sub["prev_len"])
    count = sum(dl)
    loss_sum = EO.xe_loss(preds, caps_sorted, dl) * count
    keys = list(sd)
    grads = torch.autograd.grad(loss_sum, [sd[k] for k in keys], allow_unused=True)
    flat = torch.cat([(g if g is not None else torch.zeros_like(sd[k])).reshape(-1) for k, g in zip(keys, grads)])
    return flat, count


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from show_edit_tell_b200 import parallel
    r, w = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.set_num_threads(1)
    c = CFG
    sd = EO.init_state_dict(c["V"], c["D"], c["D"], c["D"], c["A"], c["Fdim"], seed=2)
    batch = synth.make_batch(c["B"], c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True, seed=3,
                             min_len=3, min_prev=2)
    lo, hi = parallel.shard_rows(c["B"], rank, world)
    flat, count = _shard_sum_grads(sd, batch, lo, hi)
    n = flat.numel()
    buf = torch.zeros(n + 64)
    buf[:n] = flat
    slot = parallel.allreduce_sums(buf, n, torch.tensor(float(count)))
    if rank == 0:
        torch.save({"grad": buf[:n] / slot, "count": float(slot)}, out)
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process(tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    c = CFG
    sd = EO.init_state_dict(c["V"], c["D"], c["D"], c["D"], c["A"], c["Fdim"], seed=2)
    batch = synth.make_batch(c["B"], c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True, seed=3,
                             min_len=3, min_prev=2)
    flat, count = _shard_sum_grads(sd, batch, 0, c["B"])
    assert got["count"] == count
    ref = flat / count
    assert (got["grad"] - ref).abs().max() < 1e-6 * max(1.0, float(ref.abs().max()))


def test_shard_rows_covers_batch():
    from show_edit_tell_b200 import parallel
    for n, w in ((64, 8), (10, 4), (3, 2), (5, 8)):
        spans = [parallel.shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
