"""CPU restatement of the reference's homography adaptation (TEST INFRASTRUCTURE ONLY - never imported by the
product path).

Reference: homography_adaptation, python/src/homographies.py:250-324, called by
InferenceWrapper.run_with_homography_adaptation (python/src/inferencewrapper.py:48-68) and the COCO pseudo-labelling
job (python/src/preprocess_coco.py:64-74).  The random homographies (sample_homography, homographies.py:79-196) are
an INPUT here.  Third-party arithmetic restated from its published behaviour: torchvision ~0.10
functional_tensor.perspective (the sampling grid ((c0 x + c1 y + c2) / (c6 x + c7 y + 1), ...) at pixel centres x + 0.5,
then torch grid_sample(align_corners=False, zeros padding)) and OpenCV's elliptical structuring element + erode with
a constant zero border (homographies.py:239-247).  Pinned by tests/golden/homography_kat.npz.
"""
import numpy as np
import torch
import torch.nn.functional as F


def invert(h8):
    """Flattened homography [8] -> flattened inverse (homographies.py:199-203)."""
    m = np.concatenate([np.asarray(h8, np.float64), [1.0]]).reshape(3, 3)
    inv = np.linalg.inv(m)
    return (inv / inv[2, 2]).reshape(9)[:8].astype(np.float32)


def perspective(t, c, mode):
    """t [N, C, H, W] or [C, H, W]; output(x, y) = input(transform_c(x + 0.5, y + 0.5) - 0.5)."""
    squeeze = t.dim() == 3
    if squeeze:
        t = t[None]
    n, _, h, w = t.shape
    c = [float(v) for v in c]
    xs = torch.linspace(0.5, w - 0.5, w)
    ys = torch.linspace(0.5, h - 0.5, h)
    y, x = torch.meshgrid(ys, xs, indexing='ij')
    den = c[6] * x + c[7] * y + 1.0
    gx = (c[0] / (0.5 * w) * x + c[1] / (0.5 * w) * y + c[2] / (0.5 * w)) / den - 1.0
    gy = (c[3] / (0.5 * h) * x + c[4] / (0.5 * h) * y + c[5] / (0.5 * h)) / den - 1.0
    grid = torch.stack([gx, gy], -1)[None].expand(n, -1, -1, -1)
    out = F.grid_sample(t, grid, mode=mode, padding_mode='zeros', align_corners=False)
    return out[0] if squeeze else out


def ellipse(ksize):
    """cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (ksize, ksize)): row i holds ones in [c - dx, c + dx]."""
    r = c = ksize // 2
    k = np.zeros((ksize, ksize), np.uint8)
    inv_r2 = 1.0 / (r * r) if r else 0.0
    for i in range(ksize):
        dy = i - r
        if abs(dy) <= r:
            dx = int(np.rint(c * np.sqrt((r * r - dy * dy) * inv_r2)))
            k[i, max(c - dx, 0):min(c + dx + 1, ksize)] = 1
    return k


def erode(m, radius):
    """m [H, W] of 0/1 -> eroded with the (2 radius)^2 ellipse, anchor at its centre, zeros outside the image."""
    k = ellipse(2 * radius)
    a = radius
    h, w = m.shape
    pad = np.zeros((h + 2 * radius, w + 2 * radius), m.dtype)
    pad[a:a + h, a:a + w] = m
    out = np.ones_like(m)
    for i in range(2 * radius):
        for j in range(2 * radius):
            if k[i, j]:
                out = np.minimum(out, pad[i:i + h, j:j + w])
    return out


def valid_maps(hs, shape, margin):
    """Per homography: (count, mask) [H, W] float32 - where the back-projection / the warp of the image is valid."""
    ones = torch.ones((1,) + tuple(shape))
    counts, masks = [], []
    for h8 in hs:
        cnt = perspective(ones, invert(h8), 'nearest')[0].numpy()
        msk = perspective(ones, h8, 'nearest')[0].numpy()
        if margin:
            cnt, msk = erode(cnt, margin), erode(msk, margin)
        counts.append(cnt)
        masks.append(msk)
    if not counts:
        return np.zeros((0,) + tuple(shape), np.float32), np.zeros((0,) + tuple(shape), np.float32)
    return np.stack(counts), np.stack(masks)


def homography_adaptation(image, net, hs, margin=8, aggregation='sum'):
    """image [B, C, H, W] float32 tensor, net(image) -> prob [B, H, W] tensor, hs [num][8] -> prob [B, H, W]."""
    num = len(hs)
    shape = image.shape[2:4]
    probs = [net(image)]
    counts = [torch.ones_like(probs[0])]
    cnts, msks = valid_maps(hs, shape, margin)
    for k, h8 in enumerate(hs):
        warped = perspective(image, h8, 'bilinear')
        wp = net(warped) * torch.from_numpy(msks[k])[None]
        proj = perspective(wp[:, None], invert(h8), 'bilinear')[:, 0] * torch.from_numpy(cnts[k])[None]
        probs.append(proj)
        counts.append(torch.from_numpy(cnts[k])[None].expand_as(proj))
    probs = torch.stack(probs, -1)
    total = torch.stack(counts, -1).sum(-1)
    prob = probs.max(-1).values if aggregation == 'max' else probs.sum(-1) / total
    return torch.where(total >= num // 3, prob, torch.zeros_like(prob))
