"""The pseudo-labelling job of the reference (python/src/preprocess_coco.py:44-74) on this engine: homography adaptation over
batches of images, one ``<stem>.npz`` per image with the reference's two entries - ``image`` (3, H, W) float32 and ``points``
(3, N) float64 rows x, y, confidence - so that the reference's dataset reader (python/src/dataset_utils.py:6-40) loads them
unchanged.
"""
import os
from pathlib import Path

import numpy as np
import torch

from .homographies import HomographyConfig


def load_image(path, size, engine):
    """CocoPreprocessDataset.__getitem__ (preprocess_coco.py:22-35) with the device loader: RGB, /255, ratio-preserving
    bilinear resize, centre crop to ``size`` = (H, W) -> (3, H, W) CUDA tensor."""
    import cv2
    bgr = cv2.imread(str(path), cv2.IMREAD_COLOR)
    if bgr is None:
        raise FileNotFoundError(path)
    frame = torch.from_numpy(bgr.astype(np.float32) / 255.0)[None].to('cuda:%d' % engine.device)
    return engine.preprocess_f32(frame, int(size[0]), int(size[1]))[0]


def preprocess_coco_folder(image_paths, net_wrapper, homo_config, output_path, size=(240, 320), batch_size=16, rng=None):
    """preprocess_coco.py:64-74: points from the heatmap aggregated over random homographies, saved next to the image."""
    output_path = Path(output_path)
    output_path.mkdir(parents=True, exist_ok=True)
    paths = [str(p) for p in image_paths]
    written = []
    for i in range(0, len(paths), batch_size):
        chunk = paths[i:i + batch_size]
        batch = torch.stack([load_image(p, size, net_wrapper.engine) for p in chunk])
        points = net_wrapper.run_with_homography_adaptation(batch, homo_config, rng=rng)
        for j, p in enumerate(chunk):
            filename = Path(output_path, '%s.npz' % Path(p).stem)
            np.savez_compressed(filename, image=batch[j].cpu().numpy(), points=points[j])
            written.append(str(filename))
    return written


def preprocess_coco(coco_path, magic_point_path, settings, size=(240, 320)):
    """preprocess_coco.py:44-61: train2014 -> train, test2014 -> test."""
    from .inferencewrapper import InferenceWrapper
    print('Pre-process training COCO images:\n')
    print('Loading pre-trained Magic network...')
    net_wrapper = InferenceWrapper(weights_path=magic_point_path, settings=settings)
    print('Successfully loaded pre-trained network.')
    homo_config = HomographyConfig()
    homo_config.init_for_preprocess()
    for src, dst in (('train2014', 'train'), ('test2014', 'test')):
        folder = os.path.join(coco_path, src)
        if os.path.isdir(folder):
            files = sorted(str(p) for p in Path(folder).glob('*.*'))
            preprocess_coco_folder(files, net_wrapper, homo_config, Path(coco_path, dst), size)
