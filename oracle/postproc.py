"""Oracle: keypoint extraction and descriptor sampling on the CPU (TEST INFRASTRUCTURE).

Restates python/src/netutils.py:56-121 and python/src/nms.py:4-53 of the reference with numpy
and the C routine in ``nms_greedy.c``.
"""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def _nms_lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.oracle_nms_greedy.restype = ctypes.c_int
        _lib.oracle_nms_greedy.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def get_points(prob_map, conf_thresh=0.015, nms_dist=4, border_remove=4, top_k=0):
    """get_points (netutils.py:78-100) for ONE image -> (3, N) float64 rows x, y, confidence.

    threshold ``>=`` (netutils.py:59) -> greedy NMS over ALL candidates, border ones included
    (nms.py:37-44) -> descending confidence (netutils.py:92-93) -> border removal
    (netutils.py:95-99) -> optional truncation to the first ``top_k`` (not in the reference,
    which returns every survivor: top_k=0).
    """
    heat = np.ascontiguousarray(np.asarray(prob_map, dtype=np.float32))
    if heat.ndim == 3:
        assert heat.shape[0] == 1, 'one image at a time (the reference merges batch items, netutils.py:59-61)'
        heat = heat[0]
    h, w = heat.shape
    r = max(int(nms_dist), 0)
    cap = ((h + r) // (r + 1)) * ((w + r) // (r + 1)) + 1   # survivors are > r apart
    xs = np.empty(cap, np.int32)
    ys = np.empty(cap, np.int32)
    cf = np.empty(cap, np.float32)
    n = _nms_lib().oracle_nms_greedy(heat.ctypes.data, h, w, np.float32(conf_thresh), r, int(border_remove),
                                     cap, xs.ctypes.data, ys.ctypes.data, cf.ctypes.data)
    assert 0 <= n <= cap, n
    if top_k and top_k > 0:
        n = min(n, int(top_k))
    pts = np.zeros((3, n))
    pts[0], pts[1], pts[2] = xs[:n], ys[:n], cf[:n]
    return pts


def greedy_nms_python(points, img_h, img_w, dist_thresh):
    """Pure-python greedy NMS over an explicit (3, N) point list, small cases only.

    Same contract as corners_nms (nms.py:4-53): returns survivors sorted by descending
    confidence; ties by ascending pixel index.  Used to cross-check ``nms_greedy.c``.
    """
    n = points.shape[1]
    if n == 0:
        return np.zeros((3, 0))
    x = np.rint(points[0]).astype(np.int64)
    y = np.rint(points[1]).astype(np.int64)
    order = sorted(range(n), key=lambda i: (-points[2, i], y[i] * img_w + x[i]))
    alive = {(int(x[i]), int(y[i])) for i in range(n)}
    keep = []
    for i in order:
        if (int(x[i]), int(y[i])) not in alive:
            continue
        keep.append(i)
        for yy in range(y[i] - dist_thresh, y[i] + dist_thresh + 1):
            for xx in range(x[i] - dist_thresh, x[i] + dist_thresh + 1):
                alive.discard((int(xx), int(yy)))
    out = np.zeros((3, len(keep)))
    out[0], out[1], out[2] = x[keep], y[keep], points[2, keep]
    return out


def get_descriptors(points, desc_map, img_h, img_w):
    """get_descriptors (netutils.py:103-121) -> (C, N) float32 unit-norm columns.

    ``grid_sample(bilinear, zeros padding, align_corners=True)`` at
    gx = x / (W/2) - 1 (float64, then float32, netutils.py:110-115) samples the C*Hc*Wc map at
    ix = (gx + 1)/2 * (Wc - 1); the sampled vector is divided by its L2 norm without an epsilon
    (netutils.py:120).  ``desc_map`` is (1, C, Hc, Wc) or (C, Hc, Wc).
    """
    d = np.asarray(desc_map, dtype=np.float32)
    if d.ndim == 4:
        d = d[0]
    c, hc, wc = d.shape
    n = points.shape[1]
    if n == 0:
        return np.zeros((c, 0), np.float32)
    f32 = np.float32
    gx = (points[0] / (float(img_w) / 2.) - 1.).astype(f32)
    gy = (points[1] / (float(img_h) / 2.) - 1.).astype(f32)
    ix = (gx + f32(1)) / f32(2) * f32(wc - 1)
    iy = (gy + f32(1)) / f32(2) * f32(hc - 1)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    wx1 = ix - x0
    wy1 = iy - y0
    wx0 = (x0 + f32(1)) - ix
    wy0 = (y0 + f32(1)) - iy
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    out = np.zeros((c, n), f32)
    for yy, wy in ((y0, wy0), (y0 + 1, wy1)):
        for xx, wx in ((x0, wx0), (x0 + 1, wx1)):
            ok = (xx >= 0) & (xx < wc) & (yy >= 0) & (yy < hc)
            v = d[:, np.clip(yy, 0, hc - 1), np.clip(xx, 0, wc - 1)]
            out += v * (wx * wy * ok).astype(f32)[None, :]
    out /= np.linalg.norm(out, axis=0)[None, :]
    return out


def run(img, sd, conf_thresh=0.015, nms_dist=4, border_remove=4, top_k=0, descriptor_enabled=True):
    """Body of InferenceWrapper.run (python/src/inferencewrapper.py:38-46) for one image.

    ``img`` is (1, C, H, W) torch fp32.  Returns (points (3,N) f64, descriptors (128,N) f32,
    prob_map (H,W) f32 numpy).
    """
    from . import model
    prob, desc, _ = model.forward(img, sd, descriptor_enabled)
    h, w = img.shape[-2:]
    pts = get_points(prob.numpy(), conf_thresh, nms_dist, border_remove, top_k)
    dsc = get_descriptors(pts, desc.numpy(), h, w)
    return pts, dsc, prob[0].numpy()
