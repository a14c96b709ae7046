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


# ------------------------------------------------------------------------------------------------------------
# The persistent big-GEMM kernel (gemm_big_kernel): the same K-block pipeline, but the K-block count runs ACROSS
# tiles, tile ids travel from the producer to every other warp through a ring in shared memory (one writer, R
# readers, parity waits both ways), and the accumulator is double-buffered between the MMA issuer and the epilogue.
# ------------------------------------------------------------------------------------------------------------
def simulate_persistent(NS, groups, tile_nkb, seed, sched=4, n_epi=4):
    rng = random.Random(seed)
    readers = ["mma"] + ["conv%d" % g for g in range(groups)] + ["epi%d" % e for e in range(n_epi)]
    ntile = len(tile_nkb)
    p_full, q_full = [Bar() for _ in range(NS)], [Bar() for _ in range(NS)]
    conv, empty = [Bar() for _ in range(NS)], [Bar() for _ in range(NS)]
    acc_full, acc_empty = [Bar(), Bar()], [Bar(), Bar()]
    s_full, s_empty = [Bar() for _ in range(sched)], [Bar() for _ in range(sched)]
    s_arrivals = [0] * sched                        # reader arrivals of the phase in progress
    s_slot = [None] * sched                         # what the ring slot holds
    published = []                                  # tile ids in publication order (+ the end mark -1)
    inflight, landed = [], {"p": set(), "q": set()}
    converted, mma_done, drained = set(), set(), set()
    tiles_done = {"mma": 0}
    base = [0]
    for n in tile_nkb:
        base.append(base[-1] + n)
    total = base[-1]

    # ---- producer: publish tile `it` (or the end mark), then stream its K-blocks
    prod = {"it": 0, "g": 0, "stage": "publish"}

    def step_producer():
        it = prod["it"]
        if prod["stage"] == "publish":
            if it > ntile:
                return False
            s = it % sched
            if it >= sched and not s_empty[s].parity_wait_passes(((it // sched) & 1) ^ 1):
                return False
            if it >= sched:
                assert s_slot[s] == ("read", it - sched), "tile slot %d rewritten before every reader saw entry %d" % (s, it - sched)
            s_slot[s] = ("full", it)
            published.append(it if it < ntile else -1)
            s_full[s].phase += 1
            if it == ntile:
                prod["it"] = it + 1
                return True
            prod["stage"] = "load"
            return True
        g = prod["g"]
        if g == base[it + 1]:
            prod["it"], prod["stage"] = it + 1, "publish"
            return True
        s = g % NS
        if g >= NS:
            if not empty[s].parity_wait_passes(((g // NS) & 1) ^ 1):
                return False
            assert (g - NS) in mma_done, "producer refilled slot %d before the MMA of K-block %d" % (s, g - NS)
        inflight.append(("q", g)); inflight.append(("p", g))
        prod["g"] = g + 1
        return True

    def step_tma():
        if not inflight:
            return False
        kind, j = inflight.pop(rng.randrange(len(inflight)))
        landed[kind].add(j)
        (p_full if kind == "p" else q_full)[j % NS].phase += 1
        return True

    # ---- consumer side of the tile ring
    cursor = {r: 0 for r in readers}

    def next_tile(r):
        it = cursor[r]
        s = it % sched
        if not s_full[s].parity_wait_passes((it // sched) & 1):
            return None
        assert s_slot[s] in (("full", it), ("read", it)) or s_slot[s][1] == it, \
            "%s read tile entry %d out of slot %d holding %s" % (r, it, s, s_slot[s])
        assert len(published) > it, "%s passed sched_full before entry %d was published" % (r, it)
        s_arrivals[s] += 1
        if s_arrivals[s] == len(readers):
            s_arrivals[s] = 0
            s_slot[s] = ("read", it)
            s_empty[s].phase += 1
        cursor[r] = it + 1
        return published[it]

    cv = [{"tile": None, "g": 0, "g_end": 0, "stage": 0, "done": False} for _ in range(groups)]

    def step_converter(gi):
        st = cv[gi]
        if st["done"]:
            return False
        if st["tile"] is None:
            t = next_tile("conv%d" % gi)
            if t is None:
                return False
            if t < 0:
                st["done"] = True
                return True
            st["tile"] = t
            g0 = base[t]
            st["g"] = g0 + ((gi - g0) % groups)                # first K-block >= g0 with g % groups == gi
            st["g_end"] = base[t + 1]
            return True
        g = st["g"]
        if g >= st["g_end"]:
            st["tile"] = None
            return True
        s = g % NS
        if st["stage"] == 0:
            if not q_full[s].parity_wait_passes((g // NS) & 1):
                return False
            assert g in landed["q"], "group %d passed q_full for K-block %d before its tile landed" % (gi, g)
            st["stage"] = 1
            return True
        if not p_full[s].parity_wait_passes((g // NS) & 1):
            return False
        assert g in landed["p"], "group %d passed p_full for K-block %d before its tile landed" % (gi, g)
        assert g < NS or (g - NS) in mma_done, "operand slot %d overwritten while the MMA of K-block %d may still read it" % (s, g - NS)
        converted.add(g)
        conv[s].phase += 1
        st["g"], st["stage"] = g + groups, 0
        return True

    mm = {"tile": None, "i": 0, "tc": 0, "done": False}

    def step_mma():
        if mm["done"]:
            return False
        tc = mm["tc"]
        if mm["tile"] is None:
            t = next_tile("mma")
            if t is None:
                return False
            if t < 0:
                mm["done"] = True
                return True
            mm["tile"], mm["i"], mm["acc_ok"] = t, 0, False
            return True
        ab = tc & 1
        if not mm["acc_ok"]:
            if tc >= 2:
                if not acc_empty[ab].parity_wait_passes(((tc >> 1) & 1) ^ 1):
                    return False
                assert (tc - 2) in drained, "MMA reuses accumulator %d before tile #%d was drained" % (ab, tc - 2)
            mm["acc_ok"] = True
            return True
        t = mm["tile"]
        g = base[t] + mm["i"]
        if g == base[t + 1]:
            acc_full[ab].phase += 1
            tiles_done["mma"] = tc + 1
            mm["tile"], mm["tc"] = None, tc + 1
            return True
        if not conv[g % NS].parity_wait_passes((g // NS) & 1):
            return False
        assert g in converted, "MMA passed conv_bar for K-block %d before it was converted" % g
        mma_done.add(g)
        empty[g % NS].phase += 1
        mm["i"] += 1
        return True

    ep = [{"tile": None, "tc": 0, "done": False} for _ in range(n_epi)]
    epi_arrivals = [0, 0]

    def step_epilogue(e):
        st = ep[e]
        if st["done"]:
            return False
        if st["tile"] is None:
            t = next_tile("epi%d" % e)
            if t is None:
                return False
            if t < 0:
                st["done"] = True
                return True
            st["tile"] = t
            return True
        tc = st["tc"]
        ab = tc & 1
        if not acc_full[ab].parity_wait_passes((tc >> 1) & 1):
            return False
        assert tiles_done["mma"] > tc, "epilogue warp %d drains accumulator %d before tile #%d was multiplied" % (e, ab, tc)
        epi_arrivals[ab] += 1
        if epi_arrivals[ab] == n_epi:
            epi_arrivals[ab] = 0
            drained.add(tc)
            acc_empty[ab].phase += 1
        st["tile"], st["tc"] = None, tc + 1
        return True

    agents = ([step_producer, step_tma, step_mma] + [lambda g=g: step_converter(g) for g in range(groups)] +
              [lambda e=e: step_epilogue(e) for e in range(n_epi)])
    idle = 0
    while not (mm["done"] and all(c["done"] for c in cv) and all(x["done"] for x in ep)):
        if rng.choice(agents)():
            idle = 0
        else:
            idle += 1
            assert idle < 20000, "deadlock (K-blocks multiplied: %d of %d)" % (len(mma_done), total)
    assert len(mma_done) == total and len(drained) == ntile
    return True


def _big_config():
    src = open(SRC).read()
    groups = int(re.search(r"#define SET_TC_CONV_WARPS (\d+)", src).group(1)) // 4
    ns = int(re.search(r"struct BigCfg \{\s*static constexpr int kNS = (\d+);", src).group(1))
    sched = int(re.search(r"constexpr int kSched = (\d+);", src).group(1))
    epi = int(re.search(r"constexpr int kBigEpiWarps = (\d+);", src).group(1))
    return ns, groups, sched, epi


def test_persistent_kernel_protocol_never_aliases():
    ns, groups, sched, epi = _big_config()
    assert ns % groups == 0
    shapes = [[32] * 6, [2, 2, 2, 2, 2, 2, 2, 2, 2], [1, 38, 1, 3, 38, 2, 7], [5], [3, 1, 1, 1, 1, 64]]
    for tile_nkb in shapes:
        for seed in range(40):
            simulate_persistent(ns, groups, tile_nkb, seed, sched=sched, n_epi=epi)


def test_persistent_model_reproduces_parity_aliasing_across_tiles():
    # 3 operand slots with 2 converter groups: the parity aliasing of round 1 -- a group comes back to a slot two
    # phases later -- is there in the persistent form too (the K-block count runs across tile boundaries)
    with pytest.raises(AssertionError, match="passed [pq]_full"):
        for seed in range(400):
            simulate_persistent(3, 2, [5, 7, 4, 9], seed)
