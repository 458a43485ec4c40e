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


def _truncated_normal(rng, n, mean, std):
    """Normal(mean, std) restricted to mean +- 2 std (rejection sampling)."""
    out = np.empty((n,), np.float64)
    i = 0
    while i < n:
        v = rng.normal(mean, std, size=2 * (n - i) + 4) if std > 0 else np.full((n - i,), mean)
        v = v[np.abs(v - mean) <= 2 * std]
        take = min(len(v), n - i)
        out[i:i + take] = v[:take]
        i += take
    return out


def _uniform(rng, low, high):
    if low > high:
        low, high = high, low
    if low == high:
        high = low + 0.00001
    return rng.uniform(low, high)


def sample_homography(shape, config=None, rng=None, **kw):
    """A random homography between a patch of the image and the full frame, as the reference samples it: a centred
    crop of ``patch_ratio`` is perturbed in perspective, scaled, translated and rotated (each step keeping the patch
    inside the image unless ``allow_artifacts``), then the transform mapping the unit-square corners of the crop to
    the perturbed corners is solved for.  ``shape`` = (H, W).  Returns a float32 array [8]."""
    cfg = HomographyConfig() if config is None else config
    g = lambda k: kw.get(k, getattr(cfg, k))          # noqa: E731
    rng = np.random.default_rng() if rng is None else rng
    ratio = g('patch_ratio')
    margin = (1 - ratio) / 2
    src = margin + np.array([[0, 0], [0, ratio], [ratio, ratio], [ratio, 0]], np.float64)
    dst = src.copy()
    if g('perspective'):
        ax, ay = g('perspective_amplitude_x'), g('perspective_amplitude_y')
        if not g('allow_artifacts'):
            ax, ay = min(ax, margin), min(ay, margin)
        py = _truncated_normal(rng, 1, 0., ay / 2)[0]
        left = _truncated_normal(rng, 1, 0., ax / 2)[0]
        right = _truncated_normal(rng, 1, 0., ax / 2)[0]
        dst = dst + np.array([[left, py], [left, -py], [right, py], [right, -py]])
    if g('scaling'):
        n = g('n_scales')
        scales = np.concatenate([[1.], _truncated_normal(rng, n, 1, g('scaling_amplitude') / 2)])
        centre = dst.mean(0, keepdims=True)
        cand = (dst - centre)[None] * scales[:, None, None] + centre
        if g('allow_artifacts'):
            valid = np.arange(n)                                    # the reference's quirk: indices 0 .. n-1
        else:
            valid = np.nonzero(((cand >= 0.) & (cand < 1.)).sum((1, 2)))[0]
        dst = cand[valid[rng.integers(len(valid))]]
    if g('translation'):
        t_min, t_max = dst.min(0), (1. - dst).min(0)
        if g('allow_artifacts'):
            t_min, t_max = t_min + g('translation_overflow'), t_max + g('translation_overflow')
        dst = dst + np.array([[_uniform(rng, -t_min[0], t_max[0]), _uniform(rng, -t_min[1], t_max[1])]])
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
        dst = cand[valid[rng.integers(len(valid))]]
    size = np.array([shape[1], shape[0]], np.float64)             # (x, y) order
    p, q = src * size, dst * size
    a = np.zeros((8, 8))
    rhs = np.zeros((8,))
    for i in range(4):
        a[2 * i] = [p[i, 0], p[i, 1], 1, 0, 0, 0, -p[i, 0] * q[i, 0], -p[i, 1] * q[i, 0]]
        a[2 * i + 1] = [0, 0, 0, p[i, 0], p[i, 1], 1, -p[i, 0] * q[i, 1], -p[i, 1] * q[i, 1]]
        rhs[2 * i], rhs[2 * i + 1] = q[i, 0], q[i, 1]
    return np.linalg.solve(a, rhs).astype(np.float32)


def sample_homographies(shape, config, rng=None):
    """config.num homographies [num, 8] for one call of homography adaptation."""
    return np.stack([sample_homography(shape, config, rng) for _ in range(config.num)]) if config.num else np.zeros((0, 8), np.float32)
