"""Golden vectors for the retrieval evaluation from the REAL reference.

Runs only in the authoring container (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_metrics.py

Imports CLIP-ViP/src/utils/metrics.py unmodified (pure numpy), evaluates a synthetic text/video feature set the way
validate() does (run_pretrain.py:173-176, tasks/run_video_retrieval.py:155-172: simple and DSL, both directions) — including
duplicated items, which exercise compute_metrics' tie quirk — asserts oracle/metrics_oracle.py agrees bit-for-bit and stores
the features and the metrics.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("XP_REFERENCE_ROOT", "/root/reference")
sys.dont_write_bytecode = True

from oracle import metrics_oracle as O  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "CLIP-ViP/src/utils/metrics.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    rng = np.random.RandomState(3)
    n, d = 57, 64
    vis = rng.randn(n, d).astype(np.float32)
    txt = (vis + 0.8 * rng.randn(n, d)).astype(np.float32)          # correlated pairs: ranks spread over a few positions
    vis[7], txt[7] = vis[3], txt[3]                                 # an exact duplicate item  -> tied similarities
    vis[20] = vis[11]                                               # a duplicated video only  -> ties in the t2v direction
    vis /= np.linalg.norm(vis, axis=1, keepdims=True)
    txt /= np.linalg.norm(txt, axis=1, keepdims=True)

    sim = ref.cal_cossim(txt, vis)
    assert np.array_equal(sim, O.cal_cossim(txt, vis))
    out = {}
    for kind in ("simple", "DSL"):
        if kind == "DSL":
            sim_ref = sim * ref.np_softmax(sim * 100, axis=0)
            assert np.array_equal(sim_ref, O.dsl(sim, 100.0))
        else:
            sim_ref = sim
        for direction, m in (("t2v", sim_ref), ("v2t", sim_ref.T)):
            want = ref.compute_metrics(m)
            got = O.compute_metrics(m)
            assert all(float(a) == float(b) for a, b in zip(want, got)), (kind, direction, want, got)
            out[f"{kind}_{direction}"] = tuple(float(v) for v in want)
            g, e = O.rank_counts(m)
            out[f"{kind}_{direction}_greater"], out[f"{kind}_{direction}_equal"] = torch.from_numpy(g), torch.from_numpy(e)
    assert int(out["simple_t2v_equal"].max()) >= 2, "the fixture must contain ties"
    print({k: v for k, v in out.items() if isinstance(v, tuple)})
    torch.save({"txt": torch.from_numpy(txt), "vis": torch.from_numpy(vis), "sim": torch.from_numpy(sim), **out},
               os.path.join(HERE, "retrieval_metrics_n57.pt"))


if __name__ == "__main__":
    main()
