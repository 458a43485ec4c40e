"""Headless counterpart of the reference's demo loop (python/src/inference.py): the same helper names, no camera and no
window.  Frames come from image files; the frame loader, the network, the post-processing and the matcher all run on the
GPU (spb200_preprocess_f32 -> spb200_detect -> spb200_match), results are written as ``.npz``.
"""
import glob
import os

import numpy as np
import torch


def make_query_image(frame, img_size, engine=None):
    """python/src/inference.py:72-85: BGR -> RGB, ratio-preserving INTER_LINEAR resize, centre crop.  ``frame`` is the
    float32 BGR frame in [0, 1] (h, w, 3) that Camera.get_frame returns, ``img_size`` = (W, H).  With an engine the work
    runs on the device and the H*W*3 RGB image comes back as a numpy array like the reference's; the loop below keeps it
    on the device instead (``query_tensor``)."""
    return query_tensor(frame, img_size, engine)[0].permute(1, 2, 0).contiguous().cpu().numpy()


def query_tensor(frames, img_size, engine):
    """The same loader for a batch, staying on the device: (B, h, w, 3) or (h, w, 3) float32 BGR -> (B, 3, H, W) RGB CUDA."""
    f = torch.as_tensor(np.ascontiguousarray(frames, dtype=np.float32))
    if f.dim() == 3:
        f = f[None]
    return engine.preprocess_f32(f.to('cuda:%d' % engine.device), int(img_size[1]), int(img_size[0]))


def get_features(frame, net):
    """python/src/inference.py:99-104: (N, 3 + 128) rows of x, y, confidence, descriptor."""
    points, descriptors = net.run(frame)
    return np.hstack((points.T, descriptors.T))


def get_best_correspondences(stop_features, features, engine=None, nn_thresh=0.0):
    """python/src/inference.py:88-96 (cv2.BFMatcher(NORM_L2, crossCheck=True)): mutual nearest neighbours between the
    descriptors of ``features`` (query) and ``stop_features`` (train), on the GPU (spb200_match)."""
    from .netutils import _engine
    from .settings import SuperPointSettings
    e = engine or _engine(SuperPointSettings())
    dev = 'cuda:%d' % e.device
    nq, nt = len(features), len(stop_features)
    if nq == 0 or nt == 0:
        return np.zeros((0, features.shape[1] if nq else 131)), np.zeros((0,), np.int64)
    cap = max(nq, nt)
    a = torch.zeros((1, cap, 128), dtype=torch.float32, device=dev)
    b = torch.zeros((1, cap, 128), dtype=torch.float32, device=dev)
    a[0, :nq] = torch.from_numpy(np.ascontiguousarray(features[:, 3:], dtype=np.float32)).to(dev)
    b[0, :nt] = torch.from_numpy(np.ascontiguousarray(stop_features[:, 3:], dtype=np.float32)).to(dev)
    m, _ = e.match(a, torch.tensor([nq], dtype=torch.int32, device=dev), b, torch.tensor([nt], dtype=torch.int32, device=dev), nn_thresh)
    m = m[0, :nq].cpu().numpy()
    idx = np.nonzero(m >= 0)[0]
    return features[idx], m[idx].astype(np.int64)


def list_frames(path):
    """Image files of a directory (sorted), or the single file given."""
    if os.path.isdir(path):
        files = sorted(f for f in glob.glob(os.path.join(path, '*')) if f.lower().endswith(('.png', '.jpg', '.jpeg', '.bmp', '.ppm', '.pgm')))
    else:
        files = [path]
    if not files:
        raise FileNotFoundError('no image files in %s' % path)
    return files


def run_inference(opt, settings):
    """The reference's run_inference (python/src/inference.py:10-69) without camera and GUI: every image of ``opt.images``
    goes through the loader and the wrapper; ``<out>/<stem>.npz`` holds ``points`` (3, N) and ``descriptors`` (128, N) as
    InferenceWrapper.run returns them and, from the second frame on, ``matches`` (K, 2) index pairs (this frame, previous
    frame) of the mutual nearest neighbours.  Returns the number of frames written."""
    import cv2
    from .inferencewrapper import InferenceWrapper
    print('Loading pre-trained network...')
    net = InferenceWrapper(weights_path=opt.weights_path, settings=settings)
    print('Successfully loaded pre-trained network.')
    os.makedirs(opt.out, exist_ok=True)
    img_size = (opt.W, opt.H)
    prev = None
    n = 0
    for path in list_frames(opt.images):
        frame = cv2.imread(path, cv2.IMREAD_COLOR)
        if frame is None:
            print('Failed to read %s' % path)
            continue
        frame = frame.astype('float32') / 255.0                  # Camera.get_frame, python/src/camera.py:33
        x = query_tensor(frame, img_size, net.engine)
        pts, dsc = net.run_batch(x)
        features = np.hstack((pts[0].T, dsc[0].T))
        out = {'points': pts[0], 'descriptors': dsc[0]}
        if prev is not None:
            corr, idx = get_best_correspondences(prev, features, net.engine, getattr(settings, 'nn_thresh', 0.0))
            # rows of `corr` are rows of `features`: recover their indices through the (x, y) keys, which are unique after NMS
            key = {(float(r[0]), float(r[1])): i for i, r in enumerate(features)}
            out['matches'] = np.array([[key[(float(r[0]), float(r[1]))], j] for r, j in zip(corr, idx)], np.int64).reshape(-1, 2)
        np.savez_compressed(os.path.join(opt.out, os.path.splitext(os.path.basename(path))[0] + '.npz'), **out)
        prev = features
        n += 1
    return n
