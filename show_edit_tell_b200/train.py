"""Fused train steps for the EditNet path and their data-parallel form.

`XETrainer.step` is the reference's `train()` body (editnet.py:560-581) without autograd:
forward -> packed cross-entropy (loss + d logits in place) -> reverse pass into one flat
gradient buffer -> [one NCCL all-reduce of that buffer] -> global-norm clip 0.25 + Adam, all
of it CUDA kernels from libset_b200.so.  Nothing here synchronises with the host except the
`.tolist()` on caption lengths that the reference does too (editnet.py:497).

Data parallel (new; the reference has none): one process per GPU, each rank
holds the full parameters and takes a shard of the batch.  Ranks contribute loss *sums*; the
flat gradient buffer carries one extra slot holding the rank's token count, so a single
`all_reduce(SUM)` yields both the summed gradients and the global count, and the optimizer
kernel divides by it on the device.  The result equals the single-process step on the
concatenated batch.
"""
import ctypes as C
import warnings

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, ptr
from .editnet import _stream
from .parallel import allreduce_sums


class XETrainer:
    def __init__(self, decoder, lr=5e-4, max_norm=0.25, betas=(0.9, 0.999), eps=1e-8, process_group=None,
                 distributed=None, overlap=True, trace_overlap=False):
        self.decoder = decoder
        self.lr, self.max_norm, self.betas, self.eps = lr, max_norm, betas, eps
        self.distributed = dist.is_initialized() if distributed is None else distributed
        self.group = process_group
        self.overlap = overlap          # DP: all-reduce the early-final tail of the gradients under the reverse pass
        self.step_count = 0
        self._state = None
        self._comm_stream = None
        self._events = None
        self.trace_overlap = trace_overlap   # debugging: time the buckets of the overlapped all-reduce
        self._trace = None
        self.gpu_launches_last_step = 0

    def _bucket_events(self, n):
        """n CUDA events the library records as the buckets become final (created once; recording them here forces
        torch to create the underlying cudaEvent_t, whose handle the C ABI takes)"""
        if self._events is None or len(self._events) != n:
            self._events = [torch.cuda.Event(enable_timing=self.trace_overlap) for _ in range(n)]
            for ev in self._events:
                ev.record()
        return self._events

    def _allreduce_overlapped(self, grad, n, local_count, run_backward):
        """run_backward() enqueues the reverse pass.  The gradient buckets (contiguous ranges of the flat buffer in the
        order the pass finishes them: EditNetBase.BUCKET_FIELDS) are all-reduced on a side stream, each as soon as the
        library has recorded its event; only the last bucket (+ the count slot behind it) waits for the end of the
        pass.  Returns the count slot.  With `trace_overlap` the step also records when each bucket became final and
        when its all-reduce ended (`overlap_timeline()`)."""
        main = torch.cuda.current_stream()
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=grad.device)
        comm = self._comm_stream
        starts = list(self.decoder._bucket_offsets)
        count_slot = grad[n:n + 1]
        count_slot.copy_(local_count.reshape(1).to(count_slot.dtype))
        evs = self._bucket_events(len(starts) - 1)
        handles = (C.c_void_p * max(1, len(evs)))(*[ev.cuda_event for ev in evs])
        check(_lib.lib().set_backward_bucket_events(handles, len(evs)))
        tr = None
        if self.trace_overlap:
            tr = {"t0": torch.cuda.Event(enable_timing=True), "bwd_end": torch.cuda.Event(enable_timing=True),
                  "final": evs, "ar_end": [torch.cuda.Event(enable_timing=True) for _ in range(len(starts))],
                  "mb": [(b - a) * 4 / 1e6 for a, b in zip(starts, starts[1:] + [n + 64])]}
            tr["t0"].record(main)
        run_backward()                                   # (records event k when bucket k is final)
        if tr:
            tr["bwd_end"].record(main)
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        if multi:
            with torch.cuda.stream(comm):
                for k, ev in enumerate(evs):
                    comm.wait_event(ev)
                    dist.all_reduce(grad[starts[k]:starts[k + 1]], op=dist.ReduceOp.SUM, group=self.group)
                    if tr:
                        tr["ar_end"][k].record(comm)
            dist.all_reduce(grad[starts[-1]:n + 64], op=dist.ReduceOp.SUM, group=self.group)
            if tr:
                tr["ar_end"][-1].record(main)
            main.wait_stream(comm)
        self._trace = tr if multi else None
        return count_slot

    def overlap_timeline(self):
        """(trace_overlap=True, after a step + synchronize) ms since the start of the reverse pass: when each bucket
        became final, when its all-reduce ended, when the pass ended"""
        tr = self._trace
        if not tr:
            return None
        t0 = tr["t0"]
        return {"bucket_mb": tr["mb"], "final_ms": [t0.elapsed_time(e) for e in tr["final"]] + [t0.elapsed_time(tr["bwd_end"])],
                "allreduce_end_ms": [t0.elapsed_time(e) for e in tr["ar_end"]], "backward_end_ms": t0.elapsed_time(tr["bwd_end"])}

    def _ensure_state(self):
        flat = self.decoder.flatten_parameters()
        old = self._state
        if old is None or old["flat_ptr"] != flat.data_ptr():
            n = flat.numel()
            dev = flat.device
            # +64: slot n holds the token count that rides along with the gradients
            self._state = dict(flat_ptr=flat.data_ptr(), n=n,
                               grad=torch.zeros(n + 64, device=dev), m=torch.zeros(n, device=dev),
                               v=torch.zeros(n, device=dev), scratch=torch.zeros(2048, device=dev),
                               loss=torch.zeros(2, device=dev))
            if old is not None and old["n"] == n:
                # the flat parameter buffer was re-created (decoder.to(...), re-flatten): the Adam moments describe the
                # same parameters, so they move with it and step_count stays valid
                self._state["m"].copy_(old["m"])
                self._state["v"].copy_(old["v"])
            elif old is not None:
                warnings.warn("parameter layout changed: Adam state reset")
                self.step_count = 0
        return flat, self._state

    # optimizer checkpointing (the reference pickles the optimizer object, editnet.py:168-175).  The moments are saved
    # per parameter, keyed by the module's state_dict names -- the order of the flat buffer is an internal matter (it
    # follows the gradient buckets) and must not leak into checkpoints.
    def state_dict(self):
        _, st = self._ensure_state()
        dec = self.decoder
        views_m, views_v = dec._views(st["m"]), dec._views(st["v"])
        keys = [key for _, key in dec.FIELDS]
        return {"step": self.step_count,
                "exp_avg": {k: m.clone() for k, m in zip(keys, views_m)},
                "exp_avg_sq": {k: v.clone() for k, v in zip(keys, views_v)},
                "lr": self.lr, "betas": self.betas, "eps": self.eps, "max_norm": self.max_norm}

    def load_state_dict(self, sd):
        _, st = self._ensure_state()
        dec = self.decoder
        keys = [key for _, key in dec.FIELDS]
        for name, buf in (("exp_avg", st["m"]), ("exp_avg_sq", st["v"])):
            src = sd[name]
            if not isinstance(src, dict):
                raise ValueError("optimizer state holds a flat %s tensor: written before the moments were keyed by "
                                 "parameter name; its element order is unknown" % name)
            missing = [k for k in keys if k not in src]
            if missing:
                raise ValueError("optimizer state lacks %s for %s" % (name, ", ".join(missing[:4])))
            for k, view in zip(keys, dec._views(buf)):
                if tuple(src[k].shape) != tuple(view.shape):
                    raise ValueError("%s[%s] has shape %s, the parameter has %s" % (name, k, tuple(src[k].shape), tuple(view.shape)))
                view.copy_(src[k])
        self.step_count = int(sd["step"])
        self.lr, self.betas, self.eps = sd.get("lr", self.lr), tuple(sd.get("betas", self.betas)), sd.get("eps", self.eps)
        self.max_norm = sd.get("max_norm", self.max_norm)

    def step(self, image_features, encoded_captions, caption_lengths, encoded_previous_captions,
             previous_cap_length, image_mean=None, seed=None, host_lengths=None):
        """one optimisation step; returns the (device) mean loss of this rank's shard.  With `host_lengths` =
        (caption_lengths, previous_cap_length) as CPU tensors the step never synchronises with the host."""
        dec = self.decoder
        dec.train()
        flat, st = self._ensure_state()
        if image_features.shape[0] == 0:
            return self._empty_shard_step(flat, st)
        call = dec._prepare_xe(image_features, image_mean, encoded_captions, caption_lengths,
                               encoded_previous_captions, previous_cap_length, seed=seed, host_lengths=host_lengths)
        L = _lib.lib()
        s = call.shape
        # logits stay time-major inside the workspace; the loss kernel turns them into d logits in place
        check(L.set_editnet_xe_forward(
            C.byref(call.dims), C.byref(s), C.byref(dec._struct), ptr(call.feats), ptr(call.image_mean),
            ptr(call.caps), call.dec_host, ptr(call.prev), ptr(call.prev_len), call.seed, None, ptr(call.ws),
            call.ws.numel(), _stream()))
        inv = 1.0 if self.distributed else 0.0   # DP: sums, normalised by the global count later
        check(L.set_editnet_xe_loss_time_major(C.byref(call.dims), C.byref(s), ptr(call.caps), inv, ptr(st["loss"]),
                                               ptr(call.ws), call.ws.numel(), _stream()))
        grad = st["grad"]
        g = dec._struct_for(grad[:st["n"]])

        def run_backward():
            # the reverse pass initialises the gradient buffer itself (no 355 MB memset, no read of old gradients);
            # the padding between parameters and the slots behind them are never written and stay zero
            check(L.set_backward_overwrite_grads(1))
            check(L.set_editnet_xe_backward(
                C.byref(call.dims), C.byref(s), C.byref(dec._struct), C.byref(g), ptr(call.feats), ptr(call.caps),
                call.dec_host, ptr(call.prev), ptr(call.prev_len), call.seed, None, ptr(call.ws), call.ws.numel(),
                _stream()))

        count_dev = None
        if self.distributed and self.overlap:
            count_dev = self._allreduce_overlapped(grad, st["n"], st["loss"][1], run_backward)
        else:
            run_backward()
            if self.distributed:
                count_dev = allreduce_sums(grad, st["n"], st["loss"][1], self.group)
        self.step_count += 1
        check(_lib.lib().set_clip_adam(ptr(flat), ptr(grad), ptr(st["m"]), ptr(st["v"]), st["n"], self.step_count,
                                       self.lr, self.betas[0], self.betas[1], self.eps, self.max_norm, 1.0,
                                       ptr(count_dev), ptr(st["scratch"]), _stream()))
        self.last_call = call
        dec._last_call = call
        loss = st["loss"][0].clone()      # a fresh tensor: losses collected across steps must not alias one buffer
        if self.distributed:
            loss = loss / st["loss"][1]
        return loss

    def _empty_shard_step(self, flat, st):
        """a rank whose shard has no rows (shard_rows with more ranks than rows): nothing to compute, but the rank
        still contributes zeros and a zero count to the collective and applies the same update as its peers"""
        if not self.distributed:
            raise ValueError("empty batch")
        st["grad"].zero_()
        st["loss"].zero_()
        count_dev = allreduce_sums(st["grad"], st["n"], st["loss"][1], self.group)
        self.step_count += 1
        check(_lib.lib().set_clip_adam(ptr(flat), ptr(st["grad"]), ptr(st["m"]), ptr(st["v"]), st["n"], self.step_count,
                                       self.lr, self.betas[0], self.betas[1], self.eps, self.max_norm, 1.0,
                                       ptr(count_dev), ptr(st["scratch"]), _stream()))
        return st["loss"][0].clone()

    def grad_norm(self):
        return self._state["scratch"][1]

    def flat_grad(self):
        return self._state["grad"][:self._state["n"]]


class SCSTTrainer:
    """Self-critical step of editnet_rl.py:649-679 without autograd: greedy rollout (eval, no grad) ->
    sampled rollout (train mode, dropout on) -> reward -> RewardCriterion -> reverse pass -> [all-reduce]
    -> clip 0.25 + Adam.  `reward_fn(sample_seq, greedy_seq) -> (B, max_len) float tensor` stands where
    the reference calls CIDEr-D on the host (`get_self_critical_reward`, editnet_rl.py:611-646): that
    scorer (pyciderevalcap) is out of scope, so the caller supplies the rewards."""

    def __init__(self, decoder, lr=5e-5, max_norm=0.25, betas=(0.9, 0.999), eps=1e-8, process_group=None,
                 distributed=None, max_len=18):
        self.decoder = decoder
        self.lr, self.max_norm, self.betas, self.eps, self.max_len = lr, max_norm, betas, eps, max_len
        self.distributed = dist.is_initialized() if distributed is None else distributed
        self.group = process_group
        self.step_count = 0
        self._state = None

    _ensure_state = XETrainer._ensure_state

    def step(self, word_map, image_features, encoded_previous_captions, previous_cap_length, reward_fn,
             image_mean=None, seed=None):
        dec = self.decoder
        flat, st = self._ensure_state()
        L = _lib.lib()
        start, end = word_map['<start>'], word_map['<end>']
        dec.eval()      # greedy baseline, editnet_rl.py:665-668
        g = dec._prepare_rollout(encoded_previous_captions, previous_cap_length, image_features, image_mean, 0,
                                 self.max_len, start, end, keep=False)
        greedy_seq, _ = dec._rollout_raw(g)
        dec.train()     # sampled rollout with dropout, editnet_rl.py:669-671
        call = dec._prepare_rollout(encoded_previous_captions, previous_cap_length, image_features, image_mean, 1,
                                    self.max_len, start, end, seed=seed, keep=True)
        seq, slp = dec._rollout_raw(call)
        reward = reward_fn(seq, greedy_seq).to(slp.device, torch.float32).contiguous()
        dlp = torch.empty_like(slp)
        check(L.set_reward_criterion(seq.shape[0], seq.shape[1], ptr(slp), ptr(seq), ptr(reward), ptr(st["loss"]),
                                     ptr(dlp), _stream()))
        grad = st["grad"]
        if self.distributed:
            # ranks contribute sums over their rows: undo the local 1/sum(mask), carry the mask count along
            mask_sum = torch.cat([torch.ones_like(seq[:, :1]), (seq[:, :-1] > 0).long()], 1).sum().float()
            dlp.mul_(mask_sum)
        check(L.set_backward_overwrite_grads(1))     # (the reverse pass initialises the gradient buffer itself)
        dec._rollout_backward_raw(call, dlp, grad[:st["n"]])
        count_dev = None
        if self.distributed:
            count_dev = allreduce_sums(grad, st["n"], mask_sum, self.group)
        self.step_count += 1
        check(L.set_clip_adam(ptr(flat), ptr(grad), ptr(st["m"]), ptr(st["v"]), st["n"], self.step_count, self.lr,
                              self.betas[0], self.betas[1], self.eps, self.max_norm, 1.0, ptr(count_dev),
                              ptr(st["scratch"]), _stream()))
        self.last_call, self.last_seq, self.last_greedy = call, seq, greedy_seq
        dec._last_call = call
        return st["loss"][0].clone()

    state_dict = XETrainer.state_dict
    load_state_dict = XETrainer.load_state_dict
    grad_norm = XETrainer.grad_norm
    flat_grad = XETrainer.flat_grad
