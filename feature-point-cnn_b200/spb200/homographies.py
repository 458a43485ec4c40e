"""Host side of homography adaptation: the configuration object and the random homography sampler of the
reference (python/src/homographies.py:33-62, 79-196), written against numpy.  The sampled transforms go to
``Engine.homography_adaptation`` (spb200_homography_adaptation), which does everything else on the device.

A homography is the flattened 8-vector (c0 .. c7) of torchvision's ``perspective``: output pixel (x, y) reads the
input at ((c0 x + c1 y + c2) / (c6 x + c7 y + 1), (c3 x + c4 y + c5) / (c6 x + c7 y + 1)).
"""
from math import pi

import numpy as np


class HomographyConfig(object):
    """Same fields and defaults as the reference's HomographyConfig (homographies.py:33-62)."""

    def __init__(self):
        self.num = 15
        self.perspective = True
        self.scaling = True
        self.rotation = True
        self.translation = True
        self.n_scales = 5
        self.n_angles = 25
        self.scaling_amplitude = 0.1
        self.perspective_amplitude_x = 0.1
        self.perspective_amplitude_y = 0.1
        self.patch_ratio = 0.5
        self.max_angle = pi / 2
        self.allow_artifacts = False
        self.translation_overflow = 0.
        self.valid_border_margin = 8
        self.aggregation = 'sum'

    def init_for_preprocess(self):
        self.translation = self.rotation = self.scaling = self.perspective = True
        self.scaling_amplitude = 0.2
        self.perspective_amplitude_x = 0.2
        self.perspective_amplitude_y = 0.2
        self.allow_artifacts = True
        self.patch_ratio = 0.85


class NumpyDraws(object):
    """The three kinds of random draws sample_homography makes, from a numpy Generator.  Tests replay the reference's
    own draws (scipy truncnorm on numpy's global state, torch.randint, torch Uniform) through the same interface."""

    def __init__(self, rng=None):
        self.rng = np.random.default_rng() if rng is None else rng

    def truncated_normal(self, n, mean, std):
        # homographies.py:64-67 hands mean -+ 2 std to scipy's truncnorm as STANDARDISED bounds (loc 0, scale 1 are left
        # at their defaults): the draw is a standard normal restricted to [mean - 2 std, mean + 2 std] - for the small
        # amplitudes used here nearly uniform on that interval.  Followed as written.
        from scipy.stats import truncnorm
        return np.atleast_1d(truncnorm(mean - 2 * std, mean + 2 * std).rvs(n, random_state=self.rng)).astype(np.float64)

    def integer(self, high):
        return int(self.rng.integers(high))

    def uniform(self, low, high):
        return float(self.rng.uniform(low, high))


def _uniform(draws, low, high):
    """random_uniform (homographies.py:70-75)."""
    if low > high:
        low, high = high, low
    if low == high:
        high = low + 0.00001
    return draws.uniform(low, high)


def sample_homography(shape, config=None, rng=None, draws=None, **kw):
    """A random homography between a patch of the image and the full frame, as the reference samples it
    (homographies.py:79-196): a centred crop of ``patch_ratio`` is perturbed in perspective, scaled, translated and
    rotated (each step keeping the patch inside the image unless ``allow_artifacts``), then the transform mapping the
    source corners to the perturbed corners is solved for.  ``shape`` = (H, W).  Returns a float32 array [8].

    The reference starts with ``pts2 = pts1`` - ONE tensor under two names - and perturbs ``pts2`` in place until a step
    rebinds it (the scaling and rotation steps do, perspective and translation do not).  Its source corners therefore
    carry the perspective perturbation, and the translation too when scaling is off; with neither scaling nor rotation
    both names still mean the same tensor at the end and the transform is the identity.  Followed as written."""
    cfg = HomographyConfig() if config is None else config
    g = lambda k: kw.get(k, getattr(cfg, k))          # noqa: E731
    d = draws if draws is not None else NumpyDraws(rng)
    ratio = g('patch_ratio')
    margin = (1 - ratio) / 2
    src = margin + np.array([[0, 0], [0, ratio], [ratio, ratio], [ratio, 0]], np.float64)
    dst = src
    aliased = True                                    # dst is src (homographies.py:117)
    if g('perspective'):
        ax, ay = g('perspective_amplitude_x'), g('perspective_amplitude_y')
        if not g('allow_artifacts'):
            ax, ay = min(ax, margin), min(ay, margin)
        py = d.truncated_normal(1, 0., ay / 2)[0]
        left = d.truncated_normal(1, 0., ax / 2)[0]
        right = d.truncated_normal(1, 0., ax / 2)[0]
        dst = dst + np.array([[left, py], [left, -py], [right, py], [right, -py]])
        if aliased:
            src = dst                                 # in-place += on the shared tensor (homographies.py:127)
    if g('scaling'):
        n = g('n_scales')
        scales = np.concatenate([[1.], d.truncated_normal(n, 1, g('scaling_amplitude') / 2)])
        centre = dst.mean(0, keepdims=True)
        cand = (dst - centre)[None] * scales[:, None, None] + centre
        if g('allow_artifacts'):
            valid = np.arange(n)                                    # the reference's quirk: indices 0 .. n-1
        else:
            valid = np.nonzero(((cand >= 0.) & (cand < 1.)).sum((1, 2)))[0]
        dst = cand[valid[d.integer(len(valid))]]
        aliased = False                               # rebinding (homographies.py:145)
    if g('translation'):
        t_min, t_max = dst.min(0), (1. - dst).min(0)
        if g('allow_artifacts'):
            t_min, t_max = t_min + g('translation_overflow'), t_max + g('translation_overflow')
        dst = dst + np.array([[_uniform(d, -t_min[0], t_max[0]), _uniform(d, -t_min[1], t_max[1])]])
        if aliased:
            src = dst                                 # homographies.py:154 is an in-place += as well
    if g('rotation'):
        n = g('n_angles')
        angles = np.concatenate([[0.], np.linspace(-g('max_angle'), g('max_angle'), n)])
        centre = dst.mean(0, keepdims=True)
        rot = np.stack([np.cos(angles), -np.sin(angles), np.sin(angles), np.cos(angles)], 1).reshape(-1, 2, 2)
        cand = np.matmul(np.tile((dst - centre)[None], (n + 1, 1, 1)), rot) + centre
        if g('allow_artifacts'):
            valid = np.arange(n)
        else:
            valid = np.nonzero(((cand >= 0.) & (cand < 1.)).sum((1, 2)))[0]
        dst = cand[valid[d.integer(len(valid))]]
        aliased = False
    size = np.array([shape[1], shape[0]], np.float64)             # (x, y) order
    # `pts1 *= shape; pts2 *= shape` (homographies.py:177-178): a tensor still shared is scaled twice - source and target
    # stay equal, the solution below is the identity either way
    p, q = (src * size * size, dst * size * size) if aliased else (src * size, dst * size)
    a = np.zeros((8, 8))
    rhs = np.zeros((8,))
    for i in range(4):
        a[2 * i] = [p[i, 0], p[i, 1], 1, 0, 0, 0, -p[i, 0] * q[i, 0], -p[i, 1] * q[i, 0]]
        a[2 * i + 1] = [0, 0, 0, p[i, 0], p[i, 1], 1, -p[i, 0] * q[i, 1], -p[i, 1] * q[i, 1]]
        rhs[2 * i], rhs[2 * i + 1] = q[i, 0], q[i, 1]
    return np.linalg.solve(a, rhs).astype(np.float32)


def sample_homographies(shape, config, rng=None, draws=None):
    """config.num homographies [num, 8] for one call of homography adaptation."""
    d = draws if draws is not None else NumpyDraws(rng)
    return np.stack([sample_homography(shape, config, draws=d) for _ in range(config.num)]) if config.num else np.zeros((0, 8), np.float32)
