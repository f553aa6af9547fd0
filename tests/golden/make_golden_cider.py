"""Golden vectors for the CIDEr-D self-critical reward, produced by the REAL reference scorer.

Runs only in the development container (needs /root/reference):

    python tests/golden/make_golden_cider.py

Scorer cases (`cider_*.npz`): seeded random id captions are scored by the reference's own
`misc.rewards.get_self_critical_reward` (-> `CiderD.compute_score`, ciderD_scorer.py) in "corpus"
mode and with a cached document-frequency table (`--cached_tokens`), the oracle restatement
(oracle/cider.py) is asserted to agree to 1e-12, and inputs + reference outputs are stored.
tests/test_oracle_golden.py re-checks the oracle against the files on every CPU run; the GPU
tests compare the device scorer with them.
"""
from __future__ import annotations

import json
import os
import pickle
import sys
import tempfile
from collections import defaultdict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import cider as OC  # noqa: E402
from oracle import ref_loader  # noqa: E402


def random_captions(rng, n, vocab, lo, hi, width=16, full_rows=0, empty_rows=0):
    out = np.zeros((n, width), np.int64)
    for i in range(n):
        k = int(rng.integers(lo, hi + 1))
        out[i, :k] = rng.integers(1, vocab + 1, size=k)
    for i in range(min(full_rows, n)):             # no EOS inside the window
        out[i] = rng.integers(1, vocab + 1, size=width)
    for i in range(full_rows, min(full_rows + empty_rows, n)):
        out[i] = 0                                 # caption "0"
    return out


def make_case(rng, images, spi, vocab, copy_frac=0.3):
    gts = []
    for _ in range(images):
        ncap = int(rng.integers(3, 7))
        gts.append(random_captions(rng, ncap, vocab, 4, 15))
    B = images * spi
    gen = random_captions(rng, B, vocab, 3, 16, full_rows=1, empty_rows=1)
    greedy = random_captions(rng, B, vocab, 3, 16)
    # some hypotheses are (perturbed) copies of a reference, so the clipped products are non-zero
    for b in range(B):
        g = gts[b // spi]
        if rng.random() < copy_frac:
            gen[b] = g[int(rng.integers(0, len(g)))]
            if rng.random() < 0.5:
                gen[b, int(rng.integers(0, 4))] = int(rng.integers(1, vocab + 1))
        if rng.random() < copy_frac:
            greedy[b] = g[int(rng.integers(0, len(g)))]
    return gts, gen, greedy


def cached_table(rng, vocab, images):
    """A document-frequency table in the format of preprocess/prepro_ngrams.py:66-79,119-122:
    {'document_frequency': {tuple of id strings: float}, 'ref_len': number of images}."""
    df = defaultdict(float)
    for _ in range(images):
        caps = random_captions(rng, 5, vocab, 4, 15)
        grams = set()
        for c in caps:
            w = [str(x) for x in OC.caption_words(c)]
            for k in range(1, 5):
                for i in range(len(w) - k + 1):
                    grams.add(tuple(w[i:i + k]))
        for g in grams:
            df[g] += 1
    return df, images


def run_reference(R, gts, gen, greedy, cached=None):
    import torch
    R.CiderD_scorer = None
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        try:
            if cached is not None:
                os.makedirs(os.path.join(tmp, "data"))
                with open(os.path.join(tmp, "data", "synth.p"), "wb") as f:
                    pickle.dump({"document_frequency": cached[0], "ref_len": cached[1]}, f)
                os.chdir(tmp)                       # ciderD_scorer.py:70 opens data/<name>.p
                R.init_scorer("synth")
            else:
                R.init_scorer("corpus")
            data = {"gts": [g for g in gts]}
            cg, sc, gr = R.get_self_critical_reward(data, torch.from_numpy(gen), torch.from_numpy(greedy),
                                                    return_gen_scores=True)
        finally:
            os.chdir(cwd)
    return np.asarray(cg, np.float64), np.asarray(sc, np.float64), float(gr)


CASES = [
    ("cider_corpus_v30", dict(seed=1, images=4, spi=3, vocab=30, cached=False)),
    ("cider_corpus_v12_dense", dict(seed=2, images=6, spi=5, vocab=12, cached=False)),
    ("cider_cached_v40", dict(seed=3, images=5, spi=2, vocab=40, cached=True)),
    ("cider_corpus_v9487", dict(seed=4, images=8, spi=5, vocab=9487, cached=False)),
]


def main():
    ref_loader.load_reference()
    import misc.rewards as R
    for name, kw in CASES:
        rng = np.random.default_rng(kw["seed"])
        gts, gen, greedy = make_case(rng, kw["images"], kw["spi"], kw["vocab"])
        cached = cached_table(rng, kw["vocab"], 60) if kw["cached"] else None
        cg, sc, gr = run_reference(R, gts, gen, greedy, cached)
        if cached is not None:
            df = {tuple(int(x) for x in g): v for g, v in cached[0].items()}
            o_cg, o_sc, o_gr = OC.self_critical_reward(gts, gen, greedy, df, cached[1])
        else:
            o_cg, o_sc, o_gr = OC.self_critical_reward(gts, gen, greedy)
        for tag, a, b in (("gen", o_cg, cg), ("reward", o_sc, sc), ("greedy", o_gr, gr)):
            err = np.max(np.abs(np.asarray(a) - np.asarray(b)))
            assert err <= 1e-12, f"{name}.{tag}: {err}"
        assert np.count_nonzero(cg) > 0, "degenerate case: every score is zero"
        blob = {"meta": np.frombuffer(json.dumps(dict(name=name, **kw)).encode(), dtype=np.uint8),
                "gts": np.concatenate(gts, 0), "gts_off": np.cumsum([0] + [len(g) for g in gts]),
                "gen": gen, "greedy": greedy, "out.cider_gen": cg, "out.reward": sc,
                "out.cider_greedy": np.array(gr)}
        if cached is not None:
            keys, dfv = OC.doc_freq_table(df)
            blob["df_keys"], blob["df_val"], blob["ref_len"] = keys, dfv, np.array(float(cached[1]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(f"[golden] {name:28s} mean gen {cg.mean():.4f} greedy {gr:.4f} nonzero {np.count_nonzero(cg)}/{len(cg)} OK")


if __name__ == "__main__":
    main()
