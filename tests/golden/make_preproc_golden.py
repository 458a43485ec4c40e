"""Known answers of the frame loaders, produced by RUNNING THE REFERENCE'S CODE PATHS (build container only).

    python tests/golden/make_preproc_golden.py

  * u8_<k>: the C++ demo's loader, cpp/src/camera.cc:12-23 - cv::resize(frame, Size(W, H)) then cvtColor(BGR2GRAY) -
    through the same OpenCV (cv2 4.x) calls on 8-bit frames;
  * f32_<k>: the reference's own make_query_image (python/src/inference.py:72-85), imported from /root/reference with
    torchsummary / functional_tensor stubbed, followed by InferenceWrapper.prepare_input's HWC -> CHW.
Writes tests/golden/preproc_kat.npz (frames in, loaded images out).
"""
import os
import sys
import types

import numpy as np

if not hasattr(np, 'int'):
    np.int = int
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference/python')
sys.modules['torchsummary'] = types.SimpleNamespace(summary=lambda *a, **k: None)
import torchvision.transforms._functional_tensor as _ft      # noqa: E402
import torchvision.transforms as _T                          # noqa: E402
sys.modules['torchvision.transforms.functional_tensor'] = _ft
_T.functional_tensor = _ft
import cv2                                                   # noqa: E402
from src.inference import make_query_image                   # noqa: E402


def frame(rs, h, w):
    """A smooth colour gradient plus texture: interpolation of pure noise would hide off-by-one source positions."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    base = np.stack([xx / w * 255, yy / h * 255, (xx + yy) / (h + w) * 255], 2)
    return np.clip(base * 0.6 + rs.randint(0, 256, (h, w, 3)) * 0.4, 0, 255).astype(np.uint8)


def main():
    rs = np.random.RandomState(5)
    out = {}
    u8_cases = [(96, 128, 48, 64, 3), (60, 80, 112, 160, 3), (75, 101, 64, 80, 3), (60, 80, 64, 96, 1), (48, 64, 48, 64, 3),
                (135, 240, 64, 112, 3)]
    for k, (h, w, H, W, c) in enumerate(u8_cases):
        f = frame(rs, h, w)
        if c == 1:
            f = f[:, :, 0].copy()
        r = cv2.resize(f, (W, H))
        g = cv2.cvtColor(r, cv2.COLOR_BGR2GRAY) if c == 3 else r
        out['u8_%d_in' % k] = f
        out['u8_%d_out' % k] = g
    f32_cases = [(96, 128, 48, 64), (40, 56, 80, 112), (75, 101, 64, 80), (135, 240, 64, 112), (100, 80, 64, 96)]
    for k, (h, w, H, W) in enumerate(f32_cases):
        f = frame(rs, h, w).astype(np.float32) / 255.0           # Camera.get_frame (python/src/camera.py:33)
        q = make_query_image(f, (W, H))
        assert q.shape == (H, W, 3), q.shape
        out['f32_%d_in' % k] = f
        out['f32_%d_out' % k] = np.ascontiguousarray(q.transpose(2, 0, 1))
    np.savez_compressed(os.path.join(HERE, 'preproc_kat.npz'), **out)
    print('wrote', len(out) // 2, 'cases,', os.path.getsize(os.path.join(HERE, 'preproc_kat.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
