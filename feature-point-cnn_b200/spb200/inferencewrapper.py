"""Drop-in for the reference's ``InferenceWrapper`` (python/src/inferencewrapper.py:12-91).

``InferenceWrapper(weights_path, settings).run(img)`` -> ``(points (3,N) float64, descriptors (128,N)
float32)`` exactly as the reference documents, computed by one fused device pipeline
(spb200_detect): only the keypoints and their descriptors leave the GPU.
"""
import ctypes

import numpy as np
import torch

from . import ops
from .engine import Engine


class InferenceWrapper(object):
    def __init__(self, weights_path, settings):
        self.name = 'SuperPoint'
        self.settings = settings
        self.engine = Engine(getattr(settings, 'device', 0))
        try:
            # load_checkpoint_for_inference (python/src/saveutils.py:6-18), strict
            self.engine.load_checkpoint(weights_path)
            self.engine.finalize(getattr(settings, 'precision', 'fp16'))
        except Exception as e:   # the reference prints and exit(1)s (inferencewrapper.py:19-20)
            print('Failed to load checkpoint: %s (%s)' % (weights_path, e))
            raise SystemExit(1)
        self.descriptor_enabled = True
        self.weights_path = weights_path
        self._net = None

    @property
    def net(self):
        """The drop-in ``SuperPoint`` module (python/src/inferencewrapper.py:18) holding this checkpoint's 163 tensors.
        Built on first use from the same native checkpoint reader the engine used (no torch.load); it shares the wrapper's
        engine, so ``wrapper.net(image)`` and ``wrapper.run(image)`` run the same kernels on the same weights."""
        if self._net is None:
            from .superpoint import SuperPoint
            net = SuperPoint(self.settings)
            lib = self.engine._lib
            sd = net.state_dict()
            for key, dst in sd.items():
                shape = (ctypes.c_int64 * 8)()
                rank = ctypes.c_int()
                buf = np.empty((max(dst.numel(), 1),), np.float32)
                rc = lib.spb200_checkpoint_tensor(str(self.weights_path).encode(), key.encode(), ctypes.c_void_p(buf.ctypes.data),
                                                  buf.size, shape, ctypes.byref(rank))
                if rc != 0:
                    raise RuntimeError('checkpoint %s has no tensor %s' % (self.weights_path, key))
                dst.copy_(torch.from_numpy(buf[:dst.numel()]).reshape(dst.shape).to(dst.dtype))
            net.eval()
            net.adopt_engine(self.engine)
            self._net = net
        return self._net

    def trace(self, img, out_file_name):
        """python/src/inferencewrapper.py:83-91.  The weights-only file ``<out>_params.pt`` is written exactly as the
        reference writes it (state_dict keys without their first module prefix, new zipfile serialisation) - it is what
        the C++ side loads (cpp/src/superpoint.cc:27-53, and spb200_load_checkpoint).  ``<out>_script.pt`` is a TorchScript
        of the reference's PyTorch modules; this implementation has no PyTorch graph to trace (the network runs in
        libspb200.so), so no script file is written and the C++ facade rejects script files with a clear error."""
        self.prepare_input(img)                               # same input validation as the reference
        state_dict = {('.'.join(k.split('.')[1:])): v for k, v in self.net.state_dict().items()}
        torch.save(state_dict, out_file_name + '_params.pt', _use_new_zipfile_serialization=True)
        return out_file_name + '_params.pt'

    def _params(self):
        s = self.settings
        self.engine.set_params(s.confidence_thresh, s.nms_dist, s.border_remove, getattr(s, 'top_k', 0),
                               self.descriptor_enabled)

    def prepare_input(self, img):
        """python/src/inferencewrapper.py:70-81: HWC float32 RGB ndarray -> 1*3*H*W tensor; tensors pass through."""
        if not torch.is_tensor(img):
            assert img.ndim == 3
            assert img.dtype == np.float32, 'Image must be float32.'
            assert img.shape[2] == 3, 'Image must be rgb.'
            return torch.from_numpy(np.ascontiguousarray(img.transpose((2, 0, 1)))).unsqueeze(0)
        return img

    def run(self, img):
        """One image -> (points (3,N) float64 [x;y;conf] by descending confidence, descriptors (128,N) float32)."""
        pts, dsc = self.run_batch(self.prepare_input(img))
        assert len(pts) == 1, 'run() takes one image (the reference merges batch items, netutils.py:59-61)'
        return pts[0], dsc[0]

    @torch.no_grad()
    def run_batch(self, images):
        """B*C*H*W tensor -> per-image lists of points / descriptors (the reference applied per image)."""
        self._params()
        dev = 'cuda:%d' % self.engine.device
        x = images.to(dev, torch.float32)
        b, _, h, w = x.shape
        k = getattr(self.settings, 'top_k', 0)
        cap = max(self.engine.max_keypoints(h, w, self.settings.nms_dist), 1)
        if k:
            cap = min(cap, k)
        count, xy, conf, desc = torch.ops.spb200.detect(x.contiguous(), ops.register(self.engine), cap)
        count = count.cpu().numpy()
        pts, dsc = [], []
        for i in range(b):
            n = int(count[i])
            p = np.zeros((3, n))
            p[:2] = xy[i, :n].t().cpu().numpy()
            p[2] = conf[i, :n].cpu().numpy()
            pts.append(p)
            dsc.append(desc[i, :n].t().contiguous().cpu().numpy())
        return pts, dsc

    @torch.no_grad()
    def run_with_homography_adaptation(self, img, config, homographies=None, rng=None):
        """python/src/inferencewrapper.py:48-68: N*C*H*W images -> list of per-image points (3, n) from the heatmap
        aggregated over ``config.num`` random homographies (sampled here unless given as a (num, 8) array)."""
        from . import homographies as hg
        self._params()
        x = self.prepare_input(img).to('cuda:%d' % self.engine.device, torch.float32)
        b, _, h, w = x.shape
        hs = hg.sample_homographies((h, w), config, rng) if homographies is None else homographies
        prob = self.engine.homography_adaptation(x, hs, config.valid_border_margin, config.aggregation)
        cap = max(self.engine.max_keypoints(h, w, self.settings.nms_dist), 1)
        count, xy, conf = self.engine.nms(prob, cap)
        count = count.cpu().numpy()
        pts = []
        for i in range(b):
            n = int(count[i])
            p = np.zeros((3, n))
            p[:2] = xy[i, :n].t().cpu().numpy()
            p[2] = conf[i, :n].cpu().numpy()
            pts.append(p)
        return pts

    def forward(self, images):
        """The network triple (SuperPoint.forward, python/src/superpoint.py:91-115) for a B*C*H*W tensor."""
        self._params()
        return self.engine.forward(images.to('cuda:%d' % self.engine.device, torch.float32))
