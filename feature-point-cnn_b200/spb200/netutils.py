"""GPU mirrors of the reference's post-processing helpers (python/src/netutils.py:56-121).

Same names, arguments and return types: ``get_points`` returns a (3, N) float64 array of rows
x, y, confidence sorted by descending confidence; ``get_descriptors`` a (C, N) float32 array of
unit-norm columns.  The work runs in libspb200.so (spb200_nms / spb200_sample_descriptors).
"""
import numpy as np
import torch

from .engine import Engine

_engines = {}


def _engine(settings):
    dev = getattr(settings, 'device', 0)
    if dev not in _engines:
        _engines[dev] = Engine(dev)
    e = _engines[dev]
    e.set_params(settings.confidence_thresh, settings.nms_dist, settings.border_remove, getattr(settings, 'top_k', 0), True)
    return e


def restore_prob_map(prob_map, img_h, img_w, cell_size, settings=None):
    """python/src/netutils.py:64-75, same contract: ``prob_map`` is the SOFTMAXED B*65*Hc*Wc tensor (what
    SuperPoint.forward and make_prob_map_from_labels pass); the dustbin channel is dropped and the 64 channels of a cell
    become its 8x8 pixels -> B*H*W on the GPU."""
    from .settings import SuperPointSettings
    assert cell_size == 8, 'the engine is built for 8x8 cells (python/src/settings.py:7)'
    e = _engine(settings or SuperPointSettings())
    return e.restore_prob_map(torch.as_tensor(prob_map).to('cuda:%d' % e.device, torch.float32), img_h, img_w)


def heatmap_from_logits(logits, img_h, img_w, settings=None):
    """The fused form the engine uses itself: exp(l) / (sum exp(l) + 1e-5) (python/src/superpoint.py:111-112) followed by
    restore_prob_map, in one kernel, from the 65-channel LOGITS."""
    from .settings import SuperPointSettings
    e = _engine(settings or SuperPointSettings())
    return e.heatmap_from_logits(torch.as_tensor(logits).to('cuda:%d' % e.device, torch.float32), img_h, img_w)


def get_points(prob_map, img_h, img_w, settings, engine=None):
    """python/src/netutils.py:78-100 for a (1, H, W) or (H, W) heatmap."""
    e = engine or _engine(settings)
    prob = torch.as_tensor(prob_map).to('cuda:%d' % e.device, torch.float32)
    if prob.dim() == 2:
        prob = prob[None]
    assert prob.shape[0] == 1 and prob.shape[1] == img_h and prob.shape[2] == img_w
    cap = max(e.max_keypoints(img_h, img_w, settings.nms_dist), 1)
    count, xy, conf = e.nms(prob, cap)
    n = int(count[0].item())
    pts = np.zeros((3, n))
    if n:
        xy = xy[0, :n].cpu().numpy()
        pts[0] = xy[:, 0]
        pts[1] = xy[:, 1]
        pts[2] = conf[0, :n].cpu().numpy()
    return pts


def get_descriptors(points, descriptors_map, img_h, img_w, settings, engine=None):
    """python/src/netutils.py:103-121.  ``points`` are the (3, N) rows x, y, confidence that get_points returns: INTEGER pixel
    coordinates (the only kind the reference's pipeline produces).  The device kernel samples at integer keypoints through
    per-column / per-row position tables; fractional coordinates are truncated here and coordinates outside the image are
    clamped into it by the kernel."""
    c = descriptors_map.shape[1]
    n = points.shape[1]
    if n == 0:
        return np.zeros((c, 0))
    e = engine or _engine(settings)
    dev = 'cuda:%d' % e.device
    dmap = torch.as_tensor(descriptors_map).to(dev, torch.float32)
    xy = torch.from_numpy(np.ascontiguousarray(points[:2].T.astype(np.int32)))[None].to(dev)
    count = torch.tensor([n], dtype=torch.int32, device=dev)
    out = e.sample_descriptors(dmap, img_h, img_w, count, xy)
    return out[0, :n].t().contiguous().cpu().numpy()
