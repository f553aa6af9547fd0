"""TEST INFRASTRUCTURE ONLY (checker, never the product path).

numpy restatement of the reference's retrieval evaluation, cosine measure:
`eval_utils.i2t` (eval_utils.py:545-595) and `eval_utils.t2i` (:598-720).  Pinned against the
reference's own functions by tests/golden/make_golden_retrieval.py (golden retrieval_*.npz).
"""
from __future__ import annotations

import numpy as np


def _metrics(ranks):
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / len(ranks)                     # :586-590
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / len(ranks)
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / len(ranks)
    medr = np.floor(np.median(ranks)) + 1
    meanr = ranks.mean() + 1
    return (r1, r5, r10, medr, meanr)


def i2t(images, captions, npts=None):
    """Every fifth image row is a query over all captions; its rank is the best position of its
    five captions in the descending order of the scores (:556-584)."""
    if npts is None:
        npts = images.shape[0] // 5
    ranks, top1 = np.zeros(npts), np.zeros(npts)
    for index in range(npts):
        im = images[5 * index].reshape(1, images.shape[1])
        d = np.dot(im, captions.T).flatten()
        inds = np.argsort(d)[::-1]
        rank = 1e20
        for i in range(5 * index, 5 * index + 5):
            rank = min(rank, np.where(inds == i)[0][0])
        ranks[index], top1[index] = rank, inds[0]
    return _metrics(ranks), (ranks, top1)


def t2i(images, captions, npts=None, use_gen_sent=False):
    """Every caption is a query over the distinct images (:611-660)."""
    per = 1 if use_gen_sent else 5
    if npts is None:
        npts = images.shape[0] // per
    ims = np.array([images[i] for i in range(0, len(images), per)])
    ranks, top1 = np.zeros(per * npts), np.zeros(per * npts)
    for index in range(npts):
        queries = captions[per * index:per * index + per]
        d = np.dot(queries, ims.T)
        for i in range(len(d)):
            inds = np.argsort(d[i])[::-1]
            ranks[per * index + i] = np.where(inds == index)[0][0]
            top1[per * index + i] = inds[0]
    return _metrics(ranks), (ranks, top1)


def synth_embeddings(n_img, K, seed, per=5, noise=0.8):
    """l2-normalised image / caption embeddings where caption j of image i is a noisy copy of the
    image vector (so ranks are spread between 0 and a few dozen); images repeated `per` times as the
    loader-driven encoder produces them (eval_utils.py:283-412)."""
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((n_img, K)).astype(np.float32)
    images = np.repeat(base, per, 0)
    caps = images + noise * rng.standard_normal(images.shape).astype(np.float32)
    images /= np.linalg.norm(images, axis=1, keepdims=True)
    caps /= np.linalg.norm(caps, axis=1, keepdims=True)
    return images.astype(np.float32), caps.astype(np.float32)
