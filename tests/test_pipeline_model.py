"""A small executable model of the tensor-core GEMM's barrier protocol (show_edit_tell_b200/csrc/gemm_tc.cu): TMA
producer, converter groups that alternate K-blocks, MMA issuer; every wait is by phase PARITY as on the hardware.
Loads complete in random order and agents are scheduled randomly; the model fails if a parity wait ever lets an agent
through before the phase it means has really completed.  It reproduces the round-1 bug (3 Q slots with 2 converter
groups: a group revisits a slot two phases later and the parity aliases) and checks the shipped ring depths."""
import os
import random
import re

import pytest

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "show_edit_tell_b200", "csrc", "gemm_tc.cu")


class Bar:
    def __init__(self):
        self.phase = 0          # number of completed phases

    def parity_wait_passes(self, parity):
        # mbarrier.try_wait.parity: true iff the phase with that parity is not the one in progress
        return (self.phase & 1) != parity


def simulate(NP, NQ, groups, nkb, seed):
    rng = random.Random(seed)
    p_full, q_full = [Bar() for _ in range(NP)], [Bar() for _ in range(NQ)]
    conv, empty = [Bar() for _ in range(NQ)], [Bar() for _ in range(NQ)]
    inflight = []                                   # ("p"|"q", k-block) issued, not landed
    landed = {"p": set(), "q": set()}
    converted, mma_done = set(), set()
    prod = {"j": NQ, "init": False}
    conv_next = [g for g in range(groups)]          # next K-block of each group
    conv_stage = [0] * groups                       # 0: wait q, 1: wait p
    mma_next = 0

    def step_producer():
        if not prod["init"]:
            for j in range(min(nkb, NP)):
                inflight.append(("p", j))
            for j in range(min(nkb, NQ)):
                inflight.append(("q", j))
            prod["init"] = True
            return True
        j = prod["j"]
        if j >= nkb:
            return False
        s = j % NQ
        if not empty[s].parity_wait_passes(((j // NQ) & 1) ^ 1):
            return False
        assert (j - NQ) in mma_done, "producer passed empty_bar before the MMA of K-block %d" % (j - NQ)
        inflight.append(("q", j))
        jp = j - NQ + NP
        if jp < nkb:
            assert (jp - NP) in converted, "P slot refilled before K-block %d was converted" % (jp - NP)
            inflight.append(("p", jp))
        prod["j"] = j + 1
        return True

    def step_tma():
        if not inflight:
            return False
        kind, j = inflight.pop(rng.randrange(len(inflight)))     # out-of-order completion
        landed[kind].add(j)
        (p_full[j % NP] if kind == "p" else q_full[j % NQ]).phase += 1
        return True

    def step_converter(g):
        i = conv_next[g]
        if i >= nkb:
            return False
        if conv_stage[g] == 0:
            if not q_full[i % NQ].parity_wait_passes((i // NQ) & 1):
                return False
            assert i in landed["q"], "group %d passed q_full for K-block %d before its tile landed" % (g, i)
            conv_stage[g] = 1
            return True
        if not p_full[i % NP].parity_wait_passes((i // NP) & 1):
            return False
        assert i in landed["p"], "group %d passed p_full for K-block %d before its tile landed" % (g, i)
        converted.add(i)
        conv[i % NQ].phase += 1                      # (4 warp arrivals = one completed phase)
        conv_next[g] = i + groups
        conv_stage[g] = 0
        return True

    def step_mma():
        nonlocal mma_next
        i = mma_next
        if i >= nkb:
            return False
        if not conv[i % NQ].parity_wait_passes((i // NQ) & 1):
            return False
        assert i in converted, "MMA passed conv_bar for K-block %d before it was converted" % i
        mma_done.add(i)
        empty[i % NQ].phase += 1
        mma_next = i + 1
        return True

    agents = [step_producer, step_tma, step_mma] + [lambda g=g: step_converter(g) for g in range(groups)]
    idle = 0
    while mma_next < nkb:
        if rng.choice(agents)():
            idle = 0
        else:
            idle += 1
            assert idle < 10000, "deadlock at K-block %d" % mma_next
    return True


def _shipped_configs():
    src = open(SRC).read()
    nq64 = int(re.search(r"#define SET_TC_NQ64 (\d+)", src).group(1))
    np64 = int(re.search(r"#define SET_TC_NP64 (\d+)", src).group(1))
    groups = int(re.search(r"#define SET_TC_CONV_WARPS (\d+)", src).group(1)) // 4
    m = re.search(r"kNQ = TWIN \? (\d+) : \(\(QN <= 64\) \? SET_TC_NQ64 : (\d+)\);\s*static constexpr int kNP = TWIN \? (\d+) : "
                  r"\(\(QN <= 64\) \? SET_TC_NP64 : (\d+)\);", src)
    twin_nq, deep_nq, twin_np, deep_np = (int(x) for x in m.groups())
    return groups, [("QN=64", np64, nq64), ("QN=128 deep", deep_np, deep_nq), ("QN=128 twin", twin_np, twin_nq)]


def test_shipped_ring_depths_never_alias():
    groups, cfgs = _shipped_configs()
    for name, NP, NQ in cfgs:
        assert NP >= NQ and NQ % groups == 0, name
        for nkb in (1, 2, 5, 8, 16, 33):
            for seed in range(60):
                simulate(NP, NQ, groups, nkb, seed)


def test_model_reproduces_the_three_slot_bug():
    # the configuration that shipped for a few hours in round 1: 3 Q slots, 6 P slots, 2 converter groups
    with pytest.raises(AssertionError, match="passed q_full"):
        for seed in range(400):
            simulate(6, 3, 2, 16, seed)
