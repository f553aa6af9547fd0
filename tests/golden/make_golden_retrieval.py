"""Golden vectors for the retrieval evaluation, produced by the REAL reference functions
`eval_utils.i2t` / `eval_utils.t2i` (eval_utils.py:545-720).  Development container only:

    python tests/golden/make_golden_retrieval.py

Seeded embeddings (oracle/retrieval.py `synth_embeddings`) go through the reference's numpy loops;
the oracle restatement is asserted to agree exactly and the outputs are stored next to the seeds.
"""
import importlib.util
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import retrieval as OR  # noqa: E402

CASES = [("retrieval_n40_k32", dict(n_img=40, K=32, seed=11, noise=2.5)),
         ("retrieval_n64_k48_hard", dict(n_img=64, K=48, seed=12, noise=4.5))]


def main():
    ref_loader.load_reference()
    spec = importlib.util.spec_from_file_location("ref_eval_utils",
                                                  os.path.join(ref_loader.REFERENCE_ROOT, "eval_utils.py"))
    EU = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(EU)
    for name, kw in CASES:
        images, caps = OR.synth_embeddings(kw["n_img"], kw["K"], kw["seed"], noise=kw["noise"])
        data = [{"id": i, "file_path": f"img{i}.jpg"} for i in range(kw["n_img"])]
        m_i2t, (r_i2t, t_i2t) = EU.i2t(images, caps, return_ranks=True)
        with redirect_stdout(io.StringIO()):
            m_t2i, (r_t2i, t_t2i), _ = EU.t2i(images, caps, data, return_ranks=True)
            gen_caps = caps[0::5]
            m_gen, (r_gen, t_gen), _ = EU.t2i(images[0::5], gen_caps, data, return_ranks=True, useGenSent=True)
        o1, (or1, ot1) = OR.i2t(images, caps)
        o2, (or2, ot2) = OR.t2i(images, caps)
        o3, (or3, ot3) = OR.t2i(images[0::5], gen_caps, use_gen_sent=True)
        for a, b in ((m_i2t, o1), (m_t2i, o2), (m_gen, o3)):
            assert tuple(float(x) for x in a) == tuple(float(x) for x in b), (a, b)
        assert np.array_equal(r_i2t, or1) and np.array_equal(t_i2t, ot1)
        assert np.array_equal(r_t2i, or2) and np.array_equal(t_t2i, ot2)
        assert np.array_equal(r_gen, or3) and np.array_equal(t_gen, ot3)
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            meta=np.frombuffer(json.dumps(dict(name=name, **kw)).encode(), dtype=np.uint8),
                            i2t_metrics=np.array(m_i2t, np.float64), i2t_ranks=r_i2t, i2t_top1=t_i2t,
                            t2i_metrics=np.array(m_t2i, np.float64), t2i_ranks=r_t2i, t2i_top1=t_t2i,
                            gen_metrics=np.array(m_gen, np.float64), gen_ranks=r_gen, gen_top1=t_gen)
        print(f"[golden] {name:26s} i2t {m_i2t}  t2i {m_t2i}  gen {m_gen}")


if __name__ == "__main__":
    main()
