"""Command line of the inference side of the reference's python/main.py:9-99, headless.

    python -m spb200.main [--H 480 --W 640 --nms-dist 4 --conf-thresh 0.015 --nn-thresh 0.7] \
        inference --weights-path super_point.pt --images frames/ --out features/
    python -m spb200.main train --coco-path COCO --generate-points --magic-point-weights magic_point.pt

Same global options as the reference.  ``inference`` reads image files instead of a camera and writes ``.npz`` files instead of
drawing (spb200/inference.py); ``--out-file-name`` additionally exports ``<name>_params.pt`` (InferenceWrapper.trace).  Of the
``train`` modes only ``--generate-points`` (the COCO pseudo-labelling job, an inference workload) exists here: training is out
of scope.
"""
import argparse
import sys

from .settings import SuperPointSettings


def build_parser(settings):
    parser = argparse.ArgumentParser(description='spb200: SuperPoint inference on B200.')
    parser.add_argument('--H', type=int, default=480, help='Input image height.')
    parser.add_argument('--W', type=int, default=640, help='Input image width')
    parser.add_argument('--nms-dist', dest='nms_dist', type=int, default=settings.nms_dist, help='Non Maximum Suppression (NMS) distance.')
    parser.add_argument('--conf-thresh', dest='conf_thresh', type=float, default=settings.confidence_thresh, help='Detector confidence threshold.')
    parser.add_argument('--nn-thresh', dest='nn_thresh', type=float, default=settings.nn_thresh, help='Descriptor matching threshold).')
    parser.add_argument('--cuda', action='store_true', help='Accepted for compatibility: this implementation always runs on the GPU.')
    parser.add_argument('--precision', default=settings.precision,
                        help="fp16 (default) | bf16 | fp32, optionally with split-precision stages: fp16+layer1 | fp16+encoder | fp16+all")
    parser.add_argument('--top-k', dest='top_k', type=int, default=0, help='Keep the k strongest keypoints (0 = all, the reference).')
    sub = parser.add_subparsers(dest='run_mode', required=True)
    inf = sub.add_parser('inference')
    inf.add_argument('--weights-path', dest='weights_path', type=str, required=True, help='Path to pretrained weights file.')
    inf.add_argument('--images', type=str, required=True, help='Image file or directory (replaces --camid: there is no camera).')
    inf.add_argument('--out', type=str, default='features', help='Directory for the .npz files.')
    inf.add_argument('--out-file-name', dest='out_file_name', type=str, default=None, help='Filename prefix for the exported <name>_params.pt.')
    tr = sub.add_parser('train')
    tr.add_argument('--coco-path', dest='coco_path', type=str, help='Path to the coco dataset.')
    tr.add_argument('--generate-points', dest='generate_points', action='store_true', help='Generate points for the COCO dataset.')
    tr.add_argument('--magic-point-weights', dest='magic_point_weights', type=str, default='magicpoint.pth', help='Path to pretrained MagicPoint weights file.')
    return parser


def main(argv=None):
    settings = SuperPointSettings()
    opt = build_parser(settings).parse_args(argv)
    print(opt)
    settings.read_options(opt)
    settings.precision = opt.precision
    settings.top_k = opt.top_k
    if opt.run_mode == 'inference':
        from .inference import run_inference
        n = run_inference(opt, settings)
        print('%d frames written to %s' % (n, opt.out))
        if opt.out_file_name:
            import numpy as np
            from .inferencewrapper import InferenceWrapper
            w = InferenceWrapper(opt.weights_path, settings)
            print('Weights exported to', w.trace(np.zeros((opt.H, opt.W, 3), np.float32), opt.out_file_name))
        return 0
    if opt.run_mode == 'train' and opt.coco_path and opt.generate_points:
        from .preprocess_coco import preprocess_coco
        print('Pre-processing COCO dataset...')
        preprocess_coco(opt.coco_path, opt.magic_point_weights, settings)
        print('Pre-processing finished')
        return 0
    print('Only "inference" and "train --coco-path ... --generate-points" exist in this implementation (training is out of scope).')
    return 2


if __name__ == '__main__':
    sys.exit(main())
