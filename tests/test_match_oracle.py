"""Oracle of the descriptor matching step against the golden vectors produced by the reference's own
get_best_correspondences (python/src/inference.py:88-96; tests/golden/make_match_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import matching

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'match_kat.npz')


@pytest.mark.parametrize('name', ['net', 'rand', 'near', 'tiny'])
def test_oracle_matches_reference_matcher(name):
    k = np.load(GOLDEN)
    qi, ti, d = matching.mutual_nearest(k[name + '_q'], k[name + '_t'])
    assert list(qi) == list(k[name + '_qidx'])
    assert list(ti) == list(k[name + '_tidx'])
    np.testing.assert_allclose(d, k[name + '_dist'], atol=1e-6)


def test_oracle_edge_cases():
    e = np.zeros((0, 128), np.float32)
    a = np.eye(4, 128, dtype=np.float32)
    assert len(matching.mutual_nearest(e, a)[0]) == 0 and len(matching.mutual_nearest(a, e)[0]) == 0
    qi, ti, d = matching.mutual_nearest(a, a)                       # identical sets: the identity, distance 0
    assert list(qi) == [0, 1, 2, 3] and list(ti) == [0, 1, 2, 3] and float(d.max()) == 0.0
    dup = np.vstack([a[0], a[0], a[1]])                             # duplicated query: the lowest index wins the tie
    qi, ti, _ = matching.mutual_nearest(dup, a)
    assert list(qi) == [0, 2] and list(ti) == [0, 1]
    qi, _, _ = matching.mutual_nearest(a, a[::-1] * 0.5, max_dist=0.4)   # gate: distance 0.5 is not below 0.4
    assert len(qi) == 0
