"""CPU restatement of the reference's descriptor matching (TEST INFRASTRUCTURE ONLY - never imported by the
product path).

Reference: python/src/inference.py:88-96 - cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match(query, train): a query
descriptor is matched to its nearest train descriptor (L2) when that train descriptor's nearest query is the query
itself; ties go to the lowest index.  cpp/src/main.cc:9-29 accepts a correspondence whose distance is below a
tolerance; python/src/settings.py:6 (nn_thresh) is that gate.  OpenCV (4.x, not vendored by the reference) is a
third-party dependency; parity is pinned by tests/golden/match_kat.npz, produced by the reference's own function.
"""
import numpy as np


def mutual_nearest(query, train, max_dist=0.0):
    """query (Nq, D), train (Nt, D) float32 -> (queryIdx, trainIdx, distance) of the mutual nearest neighbours,
    ordered by queryIdx like BFMatcher.match."""
    q = np.asarray(query, np.float32)
    t = np.asarray(train, np.float32)
    if len(q) == 0 or len(t) == 0:
        return np.zeros((0,), np.int64), np.zeros((0,), np.int64), np.zeros((0,), np.float32)
    d2 = np.zeros((len(q), len(t)), np.float32)
    for s in range(0, len(q), 256):                       # (a - b)^2 summed in float32, in blocks
        diff = q[s:s + 256, None, :] - t[None, :, :]
        d2[s:s + 256] = np.einsum('qtd,qtd->qt', diff, diff)
    nn_q = d2.argmin(1)                                   # first minimum = lowest index on a tie
    nn_t = d2.argmin(0)
    qi = np.arange(len(q))
    keep = nn_t[nn_q] == qi
    dist = np.sqrt(d2[qi, nn_q])
    if max_dist > 0:
        keep &= dist < max_dist
    return qi[keep].astype(np.int64), nn_q[keep].astype(np.int64), dist[keep].astype(np.float32)
