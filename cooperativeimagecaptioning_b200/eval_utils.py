"""Retrieval evaluation on the device: `i2t` / `t2i` with the reference's signatures and return
values (eval_utils.py:545-595, :598-720).

The reference loops over the queries in numpy: one `np.dot` row, one `argsort` and up to five
`np.where` per query.  Here the whole score matrix is one fp32 contraction (`coopcap_retrieval_scores`,
plain FMAs: ranks must not depend on bf16 / tf32 rounding) and `coopcap_retrieval_ranks` counts,
per query, the candidates that score strictly higher than its best correct one -- the position
the reference reads off the sorted list.  Only `measure='cosine'` (the VSEFC listener's measure,
opts `vse_measure`) is on the path; the embeddings come from `model.vse.img_enc` / `txt_enc`.
The loader-driven wrappers (`encode_data`, `evalrank`, eval_utils.py:283-543) stay with the caller.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import engine as EN
from ._lib import check


def _dev(x) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise _lib.CoopcapError("retrieval evaluation runs on CUDA only (no CPU path)")
        t = t.cuda()
    return t.detach().float().contiguous()


def score_matrix(queries: torch.Tensor, cands: torch.Tensor) -> torch.Tensor:
    """queries [Q, K] . cands [N, K]^T in fp32 (np.dot of float32 arrays in the reference)."""
    Q, K = queries.shape
    N = cands.shape[0]
    out = torch.empty(Q, N, dtype=torch.float32, device=queries.device)
    check(_lib.load().coopcap_retrieval_scores(queries.data_ptr(), cands.data_ptr(), Q, N, K, out.data_ptr(),
                                               N, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def _metrics(ranks: np.ndarray):
    """(r1, r5, r10, medr, meanr) exactly as eval_utils.py:586-590 / :700-708 compute them."""
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / len(ranks)
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / len(ranks)
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / len(ranks)
    medr = np.floor(np.median(ranks)) + 1
    meanr = ranks.mean() + 1
    return (r1, r5, r10, medr, meanr)


def i2t(images, captions, npts=None, measure="cosine", return_ranks=False):
    """Images->Text (eval_utils.py:545-595).  images, captions: (5N, K); image 5i is the query,
    captions 5i..5i+4 are correct; its rank is the best of the five."""
    if measure != "cosine":
        raise NotImplementedError("only the cosine measure is on the path (VSEFCModel.py:143-146)")
    im, cap = _dev(images), _dev(captions)
    if npts is None:
        npts = im.shape[0] // 5
    q = im[0:5 * npts:5].contiguous()
    scores = score_matrix(q, cap)
    first = (torch.arange(npts, device=q.device, dtype=torch.int32) * 5).contiguous()
    ranks, top1 = EN.retrieval_ranks(scores, first, 5)
    r = ranks.cpu().numpy().astype(np.float64)
    t1 = top1.cpu().numpy().astype(np.float64)
    return (_metrics(r), (r, t1)) if return_ranks else _metrics(r)


def t2i(images, captions, images_data=None, npts=None, measure="cosine", return_ranks=False,
        useGenSent=False):
    """Text->Images (eval_utils.py:598-720).  5 ground-truth captions per image, or 1 generated
    caption per image with useGenSent; every caption is a query over the N images."""
    if measure != "cosine":
        raise NotImplementedError("only the cosine measure is on the path (VSEFCModel.py:143-146)")
    per = 1 if useGenSent else 5
    im, cap = _dev(images), _dev(captions)
    if npts is None:
        npts = im.shape[0] // per
    ims = im[0:per * npts:per].contiguous()
    q = cap[: per * npts].contiguous()
    scores = score_matrix(q, ims)
    first = (torch.arange(per * npts, device=q.device, dtype=torch.int32) // per).contiguous()
    ranks, top1 = EN.retrieval_ranks(scores, first, 1)
    r = ranks.cpu().numpy().astype(np.float64)
    t1 = top1.cpu().numpy().astype(np.float64)
    if not return_ranks:
        return _metrics(r)
    images_ranking = {}
    if images_data is not None:
        # rank of the correct image and the four best-scoring images per query (:658-697)
        top4 = torch.topk(scores, min(4, scores.shape[1]), dim=1)[1].cpu().numpy()
        for qi in range(per * npts):
            index, i = divmod(qi, per)
            rec = {"image_id": images_data[index]["id"], "rank_correct_im": r[qi],
                   "file_path": images_data[index]["file_path"]}
            for j in range(top4.shape[1]):
                rec["im_id_rank_" + str(j)] = images_data[int(top4[qi, j])]["id"]
                rec["im_url_rank_" + str(j)] = images_data[int(top4[qi, j])]["file_path"]
            if useGenSent:
                images_ranking[index] = rec
            else:
                images_ranking.setdefault(index, {})["caption" + str(i)] = rec
    return _metrics(r), (r, t1), images_ranking
