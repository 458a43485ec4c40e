"""Homography adaptation: the oracle against the golden vectors produced by the reference's own
homography_adaptation (tests/golden/make_homography_golden.py), and the host-side sampler.  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import homography as oh, model, weights

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'feature-point-cnn_b200'))


def test_ellipse_and_erosion_match_opencv_outputs():
    k = np.load(os.path.join(GOLDEN, 'homography_kat.npz'))
    assert np.array_equal(oh.ellipse(16), k['ellipse16']) and np.array_equal(oh.ellipse(6), k['ellipse6'])
    raw = oh.perspective(torch.ones((1, 240, 320)), k['default_H'][0], 'nearest')[0].numpy()
    assert np.array_equal(raw.astype(np.uint8), k['mask0_raw'])
    assert np.array_equal(oh.erode(raw, 8).astype(np.uint8), k['mask0_eroded'])


@pytest.mark.parametrize('name', ['default', 'preprocess'])
def test_oracle_matches_reference_homography_adaptation(name):
    k = np.load(os.path.join(GOLDEN, 'homography_kat.npz'))
    sd = weights.load_state_dict(os.path.join(GOLDEN, 'super_point.pt'))
    imgs = np.load(os.path.join(GOLDEN, 'images.npz'))
    x = torch.from_numpy(np.stack([imgs[str(i)] for i in k[name + '_ids']]).astype(np.float32) / 255.)[:, None]
    num, margin, agg = [int(v) for v in k[name + '_cfg']]
    p = oh.homography_adaptation(x, lambda im: model.forward(im, sd)[0], k[name + '_H'], margin, 'sum' if agg == 0 else 'max')
    # bilinear weights differ in the last bits (grid arithmetic order): 1e-4 on values below 1
    assert float(np.abs(p.numpy() - k[name + '_prob']).max()) <= 1e-4


def test_inverse_and_sampler_properties():
    from spb200 import homographies as hg
    rng = np.random.default_rng(3)
    cfg = hg.HomographyConfig()
    h, w = 240, 320
    for _ in range(50):
        c = hg.sample_homography((h, w), cfg, rng)
        assert c.shape == (8,) and c.dtype == np.float32 and np.isfinite(c).all()
        m = np.concatenate([c.astype(np.float64), [1.]]).reshape(3, 3)
        # the crop corners land near the image (the reference accepts a scale / rotation candidate as soon as ONE of
        # its coordinates is inside, homographies.py:143,171 - so 'inside' is not guaranteed, only bounded)
        for x, y in ((w / 4, h / 4), (3 * w / 4, h / 4), (w / 4, 3 * h / 4), (3 * w / 4, 3 * h / 4)):
            q = m @ np.array([x, y, 1.])
            assert q[2] > 0 and -w <= q[0] / q[2] <= 2 * w and -h <= q[1] / q[2] <= 2 * h
        inv = oh.invert(c)
        mi = np.concatenate([inv.astype(np.float64), [1.]]).reshape(3, 3)
        ident = m @ mi
        np.testing.assert_allclose(ident / ident[2, 2], np.eye(3), atol=1e-3)
    cfg.init_for_preprocess()
    hs = hg.sample_homographies((h, w), cfg, rng)
    assert hs.shape == (cfg.num, 8)
    ident = hg.sample_homography((h, w), cfg, rng, perspective=False, scaling=False, rotation=False, translation=False, patch_ratio=1.0)
    np.testing.assert_allclose(ident, [1, 0, 0, 0, 1, 0, 0, 0], atol=1e-5)


class ReferenceDraws(object):
    """The reference's own random draws, in its call order: scipy truncnorm on numpy's global state (truncated_normal,
    homographies.py:64-67), torch.randint (:144,:172) and torch's Uniform.sample (random_uniform, :70-75)."""

    def truncated_normal(self, n, mean, std):
        from scipy.stats import truncnorm
        return np.asarray(torch.tensor(truncnorm(mean - 2 * std, mean + 2 * std).rvs([n]), dtype=torch.float32), dtype=np.float64)

    def integer(self, high):
        return int(torch.randint(high=high, size=()))

    def uniform(self, low, high):
        return float(torch.distributions.uniform.Uniform(low, high).sample(()))


@pytest.mark.parametrize('name,seed', [('default', 11), ('preprocess', 12)])
def test_sampler_replays_the_reference_draw_for_draw(name, seed):
    """tests/golden/homography_kat.npz holds the homographies the reference's sample_homography returned under
    np.random.seed / torch.manual_seed(seed) (make_homography_golden.py).  Fed the same draws, the mirror must return the
    same transforms: this pins the quirks it follows - source corners that carry the perspective perturbation
    (pts2 = pts1 aliasing, homographies.py:117-178), standardised truncnorm bounds, 'valid' indices under allow_artifacts."""
    from spb200 import homographies as hg
    k = np.load(os.path.join(GOLDEN, 'homography_kat.npz'))
    cfg = hg.HomographyConfig()
    if name == 'preprocess':
        cfg.init_for_preprocess()
    cfg.num = int(k[name + '_cfg'][0])
    torch.manual_seed(seed)
    np.random.seed(seed)
    got = hg.sample_homographies((240, 320), cfg, draws=ReferenceDraws())
    want = k[name + '_H']
    assert got.shape == want.shape
    # float32 corner arithmetic in the reference, float64 here: the coefficients agree to ~1e-4 relative
    # float32 corner arithmetic in the reference, float64 here: compared by where the transforms send the image
    worst = 0.0
    for g, w_ in zip(got.astype(np.float64), want.astype(np.float64)):
        for x, y in ((0, 0), (319, 0), (0, 239), (319, 239), (160, 120)):
            pg = np.array([g[0] * x + g[1] * y + g[2], g[3] * x + g[4] * y + g[5]]) / (g[6] * x + g[7] * y + 1)
            pw = np.array([w_[0] * x + w_[1] * y + w_[2], w_[3] * x + w_[4] * y + w_[5]]) / (w_[6] * x + w_[7] * y + 1)
            worst = max(worst, float(np.abs(pg - pw).max()))
    print('[homography replay %s] largest displacement between the two transforms %.2e px' % (name, worst))
    assert worst <= 1e-2


def test_sampler_aliasing_quirks():
    from spb200 import homographies as hg
    rng = np.random.default_rng(5)
    # neither scaling nor rotation: source and target corners are the same tensor to the end -> identity
    c = hg.sample_homography((240, 320), hg.HomographyConfig(), rng, scaling=False, rotation=False)
    np.testing.assert_allclose(c, [1, 0, 0, 0, 1, 0, 0, 0], atol=1e-5)
    # perspective only + scaling: the perspective perturbation is in the source corners too, what is left is a pure
    # scaling about the perturbed centre: no projective terms
    c = hg.sample_homography((240, 320), hg.HomographyConfig(), rng, rotation=False, translation=False)
    assert abs(c[6]) < 1e-6 and abs(c[7]) < 1e-6 and abs(c[1]) < 1e-6 and abs(c[3]) < 1e-6
