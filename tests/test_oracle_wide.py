"""Pin the oracle to the wide set of reference outputs (tests/golden/forward_wide.npz): CPU only.

SuperPoint.forward, get_points and get_descriptors of the reference (python/src/superpoint.py:91-115,
python/src/netutils.py:78-121) on 16 + 16 images at 240x320, a subset at 480x640 and one frame at 1088x1920 (the whole
set is compared with the CUDA path in tests/test_gpu_wide.py), moderate and harsh checkpoints.
"""
import numpy as np
import pytest
import torch

from oracle import model, postproc, weights
from _wide import wide_cases, wide_image, wide_ref, image_matches, cell_max

CASES = ([('m', n) for n in wide_cases('m', 240)] + [('m', n) for n in wide_cases('m', 480)[::4]] +
         [('m', 'shapes1088_0')] + [('h', n) for n in wide_cases('h', 240)[::2]] + [('h', 'rand480_0')])


@pytest.fixture(scope='module')
def sds(golden_sd):
    return {'m': golden_sd, 'h': weights.make_state_dict(seed=3, preset='harsh')}


@pytest.mark.parametrize('tag,name', CASES)
def test_oracle_matches_reference_on_wide_set(tag, name, sds):
    ref = wide_ref(tag, name)
    gray = wide_image(name)
    assert image_matches(gray, ref), 'the seeded image generator no longer reproduces the fixture input'
    h, w = gray.shape
    prob, desc, logits = model.forward(gray[None, None], sds[tag])
    np.testing.assert_allclose(cell_max(prob[0].numpy(), h, w), ref['cell'], atol=2e-6)
    assert abs(prob.double().sum().item() - float(ref['hsum'])) <= 1e-6 * abs(float(ref['hsum'])) + 1e-3
    assert abs(logits.double().sum().item() - float(ref['lsum'])) <= 2e-6 * abs(float(ref['lsum'])) + 1e-2
    pts = postproc.get_points(prob[0].numpy())
    assert pts.shape[1] == ref['xy'].shape[0]
    got = {(int(x), int(y)) for x, y in zip(pts[0], pts[1])}
    want = {(int(x), int(y)) for x, y in ref['xy']}
    # the forward differs from the reference's by ~1e-7 (same ATen kernels): identical sets bar a threshold-edge point
    assert len(got ^ want) <= 2, (len(got ^ want), len(want))
    np.testing.assert_allclose(np.sort(pts[2])[::-1][:100], np.sort(ref['conf'])[::-1][:100], atol=2e-6)
    rp = np.zeros((3, min(32, ref['xy'].shape[0])))
    rp[:2] = ref['xy'][:rp.shape[1]].T
    dsc = postproc.get_descriptors(rp, desc.numpy(), h, w)
    np.testing.assert_allclose(dsc, ref['desc'][:, :rp.shape[1]], atol=2e-5)
