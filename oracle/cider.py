"""TEST INFRASTRUCTURE ONLY (checker, never the product path).

CPU float64 restatement of the CIDEr-D self-critical reward of the reference:

  * `misc/rewards.py:26-71`            array_to_str, get_self_critical_reward
  * `cider/pyciderevalcap/ciderD/ciderD_scorer.py:13-28`   precook (n-gram counts, n = 1..4)
  * `ciderD_scorer.py:105-118`         compute_doc_freq ("corpus" mode)
  * `ciderD_scorer.py:120-203`         counts2vec / sim / compute_cider
  * `ciderD.py:31-55`                  CiderD.compute_score (one entry per hypothesis)
  * `models/AlternatingJointModel.py:409-431`  traditional_cider (the loss term)

Captions are id arrays; the reference turns them into strings of decimal ids and splits them
again, so an n-gram of words is an n-gram of ids.  A caption ends WITH its first 0 (the 0 is a
word, rewards.py:29-32).  Dictionaries keep insertion order (k = 1..4 outer, position inner), which
fixes the order of every floating-point sum below exactly as in the reference.

Pinned against the reference's own scorer by tests/golden/make_golden_cider.py (golden files
tests/golden/cider_*.npz).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

N = 4
SIGMA = 6.0


def caption_words(arr: Sequence[int]) -> Tuple[int, ...]:
    """ids up to and including the first 0                                (rewards.py:26-32)"""
    out = []
    for w in arr:
        out.append(int(w))
        if int(w) == 0:
            break
    return tuple(out)


def precook(words: Sequence[int], n: int = N) -> Dict[tuple, int]:
    """n-gram -> count, insertion ordered                          (ciderD_scorer.py:13-28)"""
    counts: Dict[tuple, int] = {}
    for k in range(1, n + 1):
        for i in range(len(words) - k + 1):
            g = tuple(words[i:i + k])
            counts[g] = counts.get(g, 0) + 1
    return counts


def corpus_doc_freq(crefs: List[List[Dict[tuple, int]]]) -> Dict[tuple, float]:
    """Number of ENTRIES whose references contain the n-gram        (ciderD_scorer.py:105-118).
    An image that serves several hypotheses is counted once per hypothesis."""
    df: Dict[tuple, float] = {}
    for refs in crefs:
        for g in set(g for ref in refs for g in ref):
            df[g] = df.get(g, 0.0) + 1.0
    return df


def _counts2vec(cnts, df, log_ref_len):
    """(ciderD_scorer.py:121-145); note `length` sums the BIGRAM counts (n == 1 is the 2-gram
    slot), i.e. words - 1, a quirk kept as is."""
    vec = [dict() for _ in range(N)]
    norm = [0.0] * N
    length = 0
    for g, tf in cnts.items():
        d = np.log(max(1.0, df.get(g, 0.0)))
        n = len(g) - 1
        vec[n][g] = float(tf) * (log_ref_len - d)
        norm[n] += pow(vec[n][g], 2)
        if n == 1:
            length += tf
    return vec, [np.sqrt(x) for x in norm], length


def _sim(vh, vr, nh, nr, lh, lr):
    """(ciderD_scorer.py:147-175): clipped dot product, cosine normalisation, Gaussian length
    penalty."""
    delta = float(lh - lr)
    val = np.zeros(N)
    for n in range(N):
        for g, w in vh[n].items():
            r = vr[n].get(g, 0.0)
            val[n] += min(w, r) * r
        if nh[n] != 0 and nr[n] != 0:
            val[n] /= (nh[n] * nr[n])
        val[n] *= np.e ** (-(delta ** 2) / (2 * SIGMA ** 2))
    return val


def ciderd_scores(hyps: List[Sequence[int]], refs: List[List[Sequence[int]]],
                  doc_freq: Optional[Dict[tuple, float]] = None,
                  ref_len: Optional[float] = None) -> np.ndarray:
    """scores[i] of hypothesis i against refs[i] (already cut word tuples).  `doc_freq` / `ref_len`
    None = "corpus" mode (df from these very entries, ref_len = number of entries,
    ciderD_scorer.py:178-179,207-212); else the cached table of `--cached_tokens`
    (ciderD_scorer.py:69-73, ref_len = number of images of the corpus)."""
    ctest = [precook(h) for h in hyps]
    crefs = [[precook(r) for r in rs] for rs in refs]
    if doc_freq is None:
        doc_freq = corpus_doc_freq(crefs)
        log_ref_len = np.log(float(len(crefs)))
    else:
        log_ref_len = np.log(float(ref_len))
    out = []
    for test, rs in zip(ctest, crefs):
        vec, norm, length = _counts2vec(test, doc_freq, log_ref_len)
        score = np.zeros(N)
        for ref in rs:
            vr, nr, lr = _counts2vec(ref, doc_freq, log_ref_len)
            score += _sim(vec, vr, norm, nr, length, lr)
        s = np.mean(score)
        s /= len(rs)
        s *= 10.0
        out.append(s)
    return np.array(out)


def self_critical_reward(gts: List[np.ndarray], gen_result: np.ndarray, greedy_res: np.ndarray,
                         doc_freq=None, ref_len=None):
    """get_self_critical_reward (rewards.py:34-71).  gts[i]: int array [n_captions_i, 16] of image
    i; gen_result / greedy_res int [B, n] with B = len(gts) * seq_per_img.
    Returns (cider_gen [B], scores = gen - greedy [B], cider_greedy mean)."""
    B = gen_result.shape[0]
    spi = B // len(gts)
    hyps = [caption_words(gen_result[i]) for i in range(B)] + \
           [caption_words(greedy_res[i]) for i in range(B)]
    g = [[caption_words(c) for c in gts[i]] for i in range(len(gts))]
    refs = [g[(i % B) // spi] for i in range(2 * B)]
    s = ciderd_scores(hyps, refs, doc_freq, ref_len)
    return s[:B], s[:B] - s[B:], s[B:].mean()


def cider_loss_coef(reward: np.ndarray, gen_result: np.ndarray):
    """traditional_cider (AlternatingJointModel.py:421-427): loss_cider = sum(logp * coef) with
    coef[b, t] = -reward[b] * mask[b, t] / sum(mask), mask = [1, (w_1 > 0), ..., (w_{n-1} > 0)]
    (gen_masks[:, 1:], :385-387)."""
    B, n = gen_result.shape
    m = np.concatenate([np.ones((B, 1), np.float32), (gen_result > 0).astype(np.float32)[:, :-1]], 1)
    return (-reward.astype(np.float32))[:, None] * m / m.sum()


def doc_freq_table(doc_freq: Dict[tuple, float]):
    """The cached table as arrays: keys int64 [K, 4] (ids, -1 padded), order int [K], df float64 [K]
    -- the layout the device scorer ingests (cooperativeimagecaptioning_b200/cider.py)."""
    K = len(doc_freq)
    keys = np.full((K, N), -1, np.int64)
    df = np.zeros(K, np.float64)
    for i, (g, v) in enumerate(doc_freq.items()):
        keys[i, :len(g)] = [int(x) for x in g]
        df[i] = float(v)
    return keys, df


def _self_test():
    rng = np.random.default_rng(0)
    gts = [rng.integers(1, 20, size=(5, 16)) for _ in range(3)]
    for g in gts:
        for r in g:
            r[rng.integers(4, 16):] = 0
    gen = rng.integers(0, 20, size=(6, 16))
    gr = rng.integers(0, 20, size=(6, 16))
    a, b, c = self_critical_reward(gts, gen, gr)
    assert a.shape == (6,) and math.isfinite(c)


if __name__ == "__main__":
    _self_test()
    print("ok")
