"""Golden vectors of homography adaptation, produced by RUNNING THE REFERENCE (build container only).

    python tests/golden/make_homography_golden.py

Runs the reference's homography_adaptation (python/src/homographies.py:250-324) with the reference's SuperPoint
holding tests/golden/super_point.pt on golden images, with the random homographies recorded as they are sampled
(sample_homography, homographies.py:79-196).  torchsummary and torchvision.transforms.functional_tensor (renamed
_functional_tensor in the installed torchvision) are stubbed, SURVEY.md 8c.  Writes tests/golden/homography_kat.npz:
per case the image ids, the config, the sampled homographies [num][8] and the aggregated probability map; plus the
reference's elliptical erosion kernel and one eroded validity mask.
"""
import os
import sys
import types

import numpy as np
import torch

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
from src import homographies as hg                           # noqa: E402
from src.settings import SuperPointSettings                  # noqa: E402
from src.superpoint import SuperPoint                        # noqa: E402
from src.saveutils import load_checkpoint_for_inference      # noqa: E402


def main():
    net = SuperPoint(SuperPointSettings())
    assert load_checkpoint_for_inference(os.path.join(HERE, 'super_point.pt'), net)
    net.eval()
    imgs = np.load(os.path.join(HERE, 'images.npz'))
    recorded = []
    real_sample = hg.sample_homography

    def recording_sample(*a, **k):
        h = real_sample(*a, **k)
        recorded.append(h.clone().numpy().reshape(8))
        return h

    hg.sample_homography = recording_sample
    out = {}
    cases = {'default': (['s240_0', 's240_1'], hg.HomographyConfig(), 11),
             'preprocess': (['s240_2'], None, 12)}
    for name, (ids, cfg, seed) in cases.items():
        if cfg is None:                                      # the COCO pseudo-labelling job, preprocess_coco.py:64-74
            cfg = hg.HomographyConfig()
            cfg.init_for_preprocess()
        cfg.num = 7
        torch.manual_seed(seed)
        np.random.seed(seed)
        recorded.clear()
        x = torch.from_numpy(np.stack([imgs[i] for i in ids]).astype(np.float32) / 255.)[:, None].repeat(1, 3, 1, 1)
        with torch.no_grad():
            prob = hg.homography_adaptation(x, net, cfg)
        out[name + '_ids'] = np.array(ids)
        out[name + '_H'] = np.stack(recorded).astype(np.float32)
        out[name + '_prob'] = prob.numpy().astype(np.float32)
        out[name + '_cfg'] = np.array([cfg.num, cfg.valid_border_margin, 0 if cfg.aggregation == 'sum' else 1], np.int32)
        print(name, ids, out[name + '_H'].shape, float(prob.max()), float((prob > 0).float().mean()))
    out['ellipse16'] = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (16, 16))
    out['ellipse6'] = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (6, 6))
    h = torch.from_numpy(out['default_H'][0])[None]
    ones = torch.ones((240, 320))[None]
    m = hg.homography_transform(ones, h, interpolation='nearest')
    out['mask0_raw'] = m.numpy()[0].astype(np.uint8)
    out['mask0_eroded'] = hg.erode(m, 8).numpy()[0].astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, 'homography_kat.npz'), **out)


if __name__ == '__main__':
    main()
