"""CPU restatement of the self-critical CIDEr-D reward (test infrastructure only).

Three layers, each with its own pin:
  * n-gram counting and the document-frequency table: `precook` / `cook_refs` / `compute_doc_freq` follow
    preprocess_rl.py:7-55 of the reference -- pinned live against those functions (tests/test_ciderd_oracle.py);
  * the reward glue: `preprocess_gd`, `array_to_str`, `get_self_critical_reward` follow editnet_rl.py:587-646 --
    pinned live by running the reference's own functions (AST-extracted) with this module's scorer injected as
    `CiderD_scorer`;
  * the scorer itself, `CiderD.compute_score`: the reference calls the un-vendored, un-pinned third-party package
    `pyciderevalcap` (README.md: "clone https://github.com/ruotianluo/cider"; not installed here, no network).  Its
    published algorithm (CIDEr-D, Vedantam et al. 2015, as implemented in pyciderevalcap/ciderD/ciderD_scorer.py) is
    restated from the paper / that file's well-known structure: tf-idf n-gram vectors for n = 1..4 with
    idf = log(N) - log(max(1, df)), clipped cosine similarity min(h, r) * r / (|h| |r|), Gaussian length penalty with
    sigma = 6 on the difference of BIGRAM counts (the `if n == 1: length += term_freq` quirk), mean over n, mean over
    references, x 10.  No copy of the package is available to run against; the layer is pinned instead by known-answer
    vectors worked by hand from that definition (tests/test_ciderd_known_answer.py: derivations in its docstring), which
    both this scorer and the device kernel must reproduce.
"""
import math
from collections import OrderedDict, defaultdict

import numpy as np


def precook(s, n=4):
    """preprocess_rl.py:7-23"""
    words = s.split()
    counts = defaultdict(int)
    for k in range(1, n + 1):
        for i in range(len(words) - k + 1):
            counts[tuple(words[i:i + k])] += 1
    return counts


def cook_refs(refs, n=4):
    """preprocess_rl.py:25-33"""
    return [precook(ref, n) for ref in refs]


def compute_doc_freq(crefs):
    """preprocess_rl.py:42-55: an n-gram counts once per image (set over the image's reference captions)"""
    df = defaultdict(float)
    for refs in crefs:
        for ngram in set(ng for ref in refs for ng in ref.keys()):
            df[ngram] += 1
    return df


class CiderD:
    """compute_score(gts, res) of pyciderevalcap's CiderD with a pre-computed document-frequency table
    (`df='coco-train-idxs'` mode of editnet_rl.py:584: {'document_frequency': {ngram tuple of str: count}, 'ref_len': N})."""

    def __init__(self, document_frequency, ref_len, n=4, sigma=6.0):
        self.df = document_frequency
        self.ref_len = math.log(float(ref_len))
        self.n, self.sigma = n, sigma

    def _counts2vec(self, cnts):
        vec = [defaultdict(float) for _ in range(self.n)]
        length = 0
        norm = [0.0] * self.n
        for ngram, tf in cnts.items():
            df = math.log(max(1.0, self.df.get(ngram, 0.0)))
            k = len(ngram) - 1
            vec[k][ngram] = float(tf) * (self.ref_len - df)
            norm[k] += vec[k][ngram] ** 2
            if k == 1:
                length += tf
        return vec, [math.sqrt(x) for x in norm], length

    def _sim(self, vh, vr, nh, nr, lh, lr):
        delta = float(lh - lr)
        val = np.zeros(self.n)
        for k in range(self.n):
            for ngram, w in vh[k].items():
                val[k] += min(w, vr[k][ngram]) * vr[k][ngram] if ngram in vr[k] else 0.0
            if nh[k] != 0 and nr[k] != 0:
                val[k] /= nh[k] * nr[k]
            val[k] *= math.e ** (-(delta ** 2) / (2 * self.sigma ** 2))
        return val

    def compute_score(self, gts, res):
        scores = []
        for r in res:
            hyp = precook(r["caption"][0], self.n)
            refs = cook_refs(gts[r["image_id"]], self.n)
            vh, nh, lh = self._counts2vec(hyp)
            score = np.zeros(self.n)
            for ref in refs:
                vr, nr, lr = self._counts2vec(ref)
                score += self._sim(vh, vr, nh, nr, lh, lr)
            s = float(np.mean(score)) / len(refs) * 10.0
            scores.append(s)
        return float(np.mean(scores)), np.array(scores)


def preprocess_gd(allcaps, word_map):
    """editnet_rl.py:587-600: drop <start>/<pad>, <end> -> 0"""
    out = []
    for j in range(allcaps.shape[0]):
        caps = allcaps[j].tolist()
        caps = [[w for w in c if w not in {word_map["<start>"], word_map["<pad>"]}] for c in caps]
        out.append([[w if w != word_map["<end>"] else 0 for w in c] for c in caps])
    return out


def array_to_str(arr):
    """editnet_rl.py:602-609: tokens up to and INCLUDING the first 0"""
    out = ""
    for x in arr:
        out += str(int(x)) + " "
        if x == 0:
            break
    return out.strip()


def self_critical_reward(scorer, gen_result, greedy_res, ground_truth, cider_weight=1):
    """editnet_rl.py:611-646 -> (B, max_len) float32 numpy array"""
    gen = np.asarray(gen_result)
    gre = np.asarray(greedy_res)
    B = gen.shape[0]
    res = OrderedDict()
    for i in range(B):
        res[i] = [array_to_str(gen[i])]
    for i in range(B):
        res[B + i] = [array_to_str(gre[i])]
    gts = {i: [array_to_str(c) for c in ground_truth[i]] for i in range(len(ground_truth))}
    res_ = [{"image_id": i, "caption": res[i]} for i in range(2 * B)]
    gts = {i: gts[i % B] for i in range(2 * B)}
    _, sc = scorer.compute_score(gts, res_)
    sc = cider_weight * sc
    sc = sc[:B] - sc[B:]
    return np.repeat(sc[:, np.newaxis], gen.shape[1], 1).astype(np.float32)


def synthetic_table(V, n_images=300, seed=0, max_len=16):
    """a small document-frequency table in the 'coco-train-idxs' format (keys: tuples of index STRINGS; <end> = 0,
    preprocess_rl.py:58-81), from random captions -- for tests and the bench"""
    rng = np.random.RandomState(seed)
    refs = []
    for _ in range(n_images):
        caps = []
        for _ in range(5):
            L = int(rng.randint(4, max_len))
            # a Zipf-like vocabulary so that n-grams repeat across images
            toks = np.minimum(rng.zipf(1.3, size=L), V - 4).astype(int)
            caps.append(" ".join(str(int(t)) for t in toks) + " 0")
        refs.append(caps)
    crefs = [cook_refs(r) for r in refs]
    return {"document_frequency": dict(compute_doc_freq(crefs)), "ref_len": n_images}
