"""Golden vectors of the descriptor matching step, produced by RUNNING THE REFERENCE (build container only).

    python tests/golden/make_match_golden.py

Imports get_best_correspondences from /root/reference/python/src/inference.py:88-96 (cv2.BFMatcher(NORM_L2,
crossCheck=True)); the two modules that file pulls in and that no longer import under the installed
torchvision / without torchsummary are stubbed.  Writes tests/golden/match_kat.npz: per case the two
descriptor sets and the reference's (queryIdx, trainIdx, distance) triples.
"""
import os
import sys
import types

import numpy as np

if not hasattr(np, 'int'):
    np.int = int
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference/python')
sys.modules['torchsummary'] = types.SimpleNamespace(summary=lambda *a, **k: None)
import torchvision.transforms._functional_tensor as _ft      # noqa: E402
import torchvision.transforms as _T                          # noqa: E402
sys.modules['torchvision.transforms.functional_tensor'] = _ft
_T.functional_tensor = _ft
import cv2                                                   # noqa: E402
from src.inference import get_best_correspondences           # noqa: E402


def unit(x):
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def reference_matches(query, train):
    """query (Nq, D), train (Nt, D) -> queryIdx, trainIdx, distance as the reference's matcher returns them."""
    fq = np.hstack([np.arange(len(query), dtype=np.float32)[:, None], np.zeros((len(query), 2), np.float32), query])
    ft = np.hstack([np.arange(len(train), dtype=np.float32)[:, None], np.zeros((len(train), 2), np.float32), train])
    corr, tidx = get_best_correspondences(ft, fq)            # (stop_features, features)
    qidx = corr[:, 0].astype(np.int64) if len(corr) else np.zeros((0,), np.int64)
    bf = cv2.BFMatcher(cv2.NORM_L2, crossCheck=True)
    ms = bf.match(queryDescriptors=query, trainDescriptors=train)
    assert [m.queryIdx for m in ms] == list(qidx) and [m.trainIdx for m in ms] == list(tidx)
    return qidx, np.asarray(tidx, np.int64), np.array([m.distance for m in ms], np.float32)


def main():
    rng = np.random.RandomState(7)
    cases = {}
    f0 = np.load(os.path.join(HERE, 'forward_shapes240_0.npz'))['descriptors'].T.astype(np.float32)
    f1 = np.load(os.path.join(HERE, 'forward_shapes240_1.npz'))['descriptors'].T.astype(np.float32)
    cases['net'] = (np.ascontiguousarray(f0), np.ascontiguousarray(f1))
    cases['rand'] = (unit(rng.randn(300, 128)), unit(rng.randn(257, 128)))
    a = unit(rng.randn(500, 128))
    perm = rng.permutation(500)[:400]
    b = unit(np.vstack([a[perm] + 0.05 * rng.randn(400, 128), rng.randn(90, 128)]))
    cases['near'] = (a, b)
    cases['tiny'] = (unit(rng.randn(1, 128)), unit(rng.randn(3, 128)))
    out = {}
    for name, (q, t) in cases.items():
        qi, ti, d = reference_matches(q, t)
        out[name + '_q'] = q; out[name + '_t'] = t
        out[name + '_qidx'] = qi; out[name + '_tidx'] = ti; out[name + '_dist'] = d
        print(name, q.shape, t.shape, len(qi), 'matches')
    np.savez_compressed(os.path.join(HERE, 'match_kat.npz'), **out)


if __name__ == '__main__':
    main()
