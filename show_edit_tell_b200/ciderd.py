"""Self-critical CIDEr-D reward on the device: `get_self_critical_reward` of editnet_rl.py:611-646 (with preprocess_gd
:587-600 and the CiderD scorer it calls) as one library call on token ids -- no GPU -> CPU -> strings -> GPU round trip.

    table = CiderDTable.from_pickle('data/coco-train-idxs.p', device)      # written by preprocess_rl.py:95-96
    reward = self_critical_reward(seq_gen, greedy_res, allcaps, word_map, table)        # (B, max_len) float32, on device

`SCSTTrainer.step(..., reward_fn=lambda s, g: self_critical_reward(s, g, allcaps, word_map, table))` closes the loop.
"""
import ctypes as C
import pickle

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr
from .editnet import _stream

_EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x):
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def pack_ngram(tokens):
    """sum_j (token_j + 1) << (16 j): exact for up to 4 tokens below 65535"""
    key = 0
    for j, t in enumerate(tokens):
        t = int(t)
        if not 0 <= t < 65535:
            raise ValueError("token id %d does not fit the 16-bit n-gram packing" % t)
        key |= (t + 1) << (16 * j)
    return key


class CiderDTable:
    """document frequencies of n-grams (n <= 4) of token ids in an open-addressing table on the device"""

    def __init__(self, document_frequency, ref_len, device):
        items = [(pack_ngram(ng), float(c)) for ng, c in document_frequency.items() if 1 <= len(ng) <= 4]
        n = len(items)
        cap = 16
        while cap < 2 * max(n, 1):
            cap *= 2
        keys = np.array([k for k, _ in items], dtype=np.uint64)
        vals = np.array([v for _, v in items], dtype=np.float32)
        tk = np.full(cap, _EMPTY, dtype=np.uint64)
        tv = np.zeros(cap, dtype=np.float32)
        mask = np.uint64(cap - 1)
        slots = _mix64(keys) & mask
        pending = np.arange(n)
        while pending.size:                                  # batched linear-probing insertion
            s = slots[pending]
            free = tk[s] == _EMPTY
            cand, cs = pending[free], s[free]
            uniq, first = np.unique(cs, return_index=True)
            winners = cand[first]
            tk[uniq] = keys[winners]
            tv[uniq] = vals[winners]
            placed = np.zeros(n, dtype=bool)
            placed[winners] = True
            pending = pending[~placed[pending]]
            slots[pending] = (slots[pending] + np.uint64(1)) & mask
        self.capacity = cap
        self.ref_len = float(ref_len)
        self.keys = torch.from_numpy(tk.view(np.int64)).to(device)      # bit pattern of the uint64 keys
        self.vals = torch.from_numpy(tv).to(device)

    @classmethod
    def from_pickle(cls, path, device):
        """the `coco-train-idxs.p` file of preprocess_rl.py:95-96: n-gram tuples of index STRINGS ('<end>' is '0')"""
        with open(path, "rb") as f:
            d = pickle.load(f, encoding="latin1")
        df = {tuple(int(w) for w in ng): c for ng, c in d["document_frequency"].items()}
        return cls(df, d["ref_len"], device)


def self_critical_reward(gen_result, greedy_res, all_captions, word_map, table, cider_weight=1.0, sigma=6.0,
                         return_scores=False):
    """gen_result, greedy_res (B, max_len) int64 rollouts; all_captions (B, 5, cap_len) int64 as the loader delivers them
    (`allcaps`, editnet_rl.py:660-673) -> rewards (B, max_len) float32 on the device"""
    if not gen_result.is_cuda:
        raise RuntimeError("show_edit_tell_b200 has no CPU path: tensors must be on a CUDA device")
    gen, gre, caps = gen_result.contiguous(), greedy_res.contiguous(), all_captions.contiguous()
    B, L = gen.shape
    R, Wc = caps.shape[1], caps.shape[2]
    scores = torch.empty(2 * B, device=gen.device)
    rewards = torch.empty(B, L, device=gen.device)
    check(_lib.lib().set_ciderd_reward(B, L, R, Wc, ptr(gen), ptr(gre), ptr(caps), word_map["<start>"], word_map["<end>"],
                                       word_map["<pad>"], ptr(table.keys), ptr(table.vals), table.capacity, table.ref_len,
                                       sigma, cider_weight, ptr(scores), ptr(rewards), _stream()))
    return (rewards, scores) if return_scores else rewards
