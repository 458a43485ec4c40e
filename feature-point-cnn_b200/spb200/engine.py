"""Thin Python owner of a spb200 engine handle; all compute happens in libspb200.so."""
import ctypes

import numpy as np
import torch

from . import _lib

PRECISIONS = {'fp32': 0, 'fp16': 1, 'bf16': 2}
# split-precision stages (include/spb200.h SPB200_SPLIT_*): a precision string may carry the level as a suffix
SPLIT_LEVELS = {'none': 0, 'layer1': 1, 'layer2': 2, 'encoder': 2, 'detector': 3, 'all': 3}

# activation buffer ids (csrc/engine.h BufId) for spb200_export_activation
BUFFERS = ['pool', 'l1a_y', 'l1a', 'l1b_y', 'l1b', 'l2a_y', 'l2a', 'l2b_y', 'feat', 'd0_y', 'd0', 'd1_y', 'logits',
           'i0_y', 'i0', 'i1_y', 'i1', 'up', 'o0_y', 'o0', 'o1_y', 'desc', 'feat_hi']


class Spb200Error(RuntimeError):
    pass


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class Engine:
    def __init__(self, device=0):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        self.device = int(device)
        rc = self._lib.spb200_create(self.device, ctypes.byref(self._h))
        if rc != 0:
            msg = self._lib.spb200_last_error(None)
            self._h = None
            raise Spb200Error('spb200_create failed: %s' % (msg.decode() if msg else rc))
        self.precision = None

    def close(self):
        if getattr(self, '_h', None):
            self._lib.spb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.spb200_last_error(self._h)
            err = Spb200Error('%s failed: %s' % (what, msg.decode() if msg else rc))
            if rc == 1:
                raise ValueError(str(err))
            raise err

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- weights -------------------------------------------------------------------------------
    def load_checkpoint(self, path):
        self._check(self._lib.spb200_load_checkpoint(self._h, str(path).encode()), 'spb200_load_checkpoint')

    def load_state_dict(self, sd):
        for k, v in sd.items():
            a = np.ascontiguousarray(v.detach().cpu().to(torch.float32).numpy())
            shape = (ctypes.c_int64 * max(a.ndim, 1))(*a.shape)
            self._check(self._lib.spb200_load_tensor(self._h, k.encode(), ctypes.c_void_p(a.ctypes.data), shape, a.ndim),
                        'spb200_load_tensor(%s)' % k)

    def finalize(self, precision='fp16', split=None):
        """precision: 'fp32' | 'fp16' | 'bf16', optionally with the split level as a suffix ('fp16+all', 'fp16+encoder',
        'fp16+layer1'); split: the same level as a separate argument (name or 0..3)."""
        if '+' in precision:
            precision, split = precision.split('+', 1)
        level = SPLIT_LEVELS[split] if isinstance(split, str) else int(split or 0)
        if level:
            self._check(self._lib.spb200_finalize_weights_split(self._h, PRECISIONS[precision], level),
                        'spb200_finalize_weights_split')
        else:
            self._check(self._lib.spb200_finalize_weights(self._h, PRECISIONS[precision]), 'spb200_finalize_weights')
        self.precision = precision
        self.split_level = level

    def set_params(self, conf_thresh=0.015, nms_dist=4, border_remove=4, top_k=0, descriptor_enabled=True):
        self._check(self._lib.spb200_set_params(self._h, float(conf_thresh), int(nms_dist), int(border_remove),
                                                int(top_k), int(bool(descriptor_enabled))), 'spb200_set_params')

    # ---- inference -----------------------------------------------------------------------------
    def _img(self, img):
        if img.dim() != 4 or img.dtype != torch.float32 or not img.is_cuda:
            raise ValueError('img must be a float32 CUDA tensor B*C*H*W')
        return img.contiguous()

    def forward(self, img, want_desc=True, want_logits=True):
        img = self._img(img)
        b, c, h, w = img.shape
        dev = img.device
        prob = torch.empty((b, h, w), dtype=torch.float32, device=dev)
        desc = torch.empty((b, 128, h // 8, w // 8), dtype=torch.float32, device=dev) if want_desc else None
        logits = torch.empty((b, 65, h // 8, w // 8), dtype=torch.float32, device=dev) if want_logits else None
        self._check(self._lib.spb200_forward(self._h, _ptr(img), b, c, h, w, _ptr(prob), _ptr(desc), _ptr(logits),
                                             self._stream()), 'spb200_forward')
        return prob, desc, logits

    def max_keypoints(self, h, w, nms_dist=4):
        return self._lib.spb200_max_keypoints(h, w, nms_dist)

    def detect(self, img, capacity, want_desc=True, want_prob=False, out=None):
        """Returns (count[B] i32, xy[B,cap,2] i32, conf[B,cap] f32, desc[B,cap,128] f32 or None, prob or None)."""
        img = self._img(img)
        b, c, h, w = img.shape
        dev = img.device
        if out is None:
            out = self.alloc_outputs(b, capacity, dev, want_desc, (h, w) if want_prob else None)
        count, xy, conf, desc, prob = out
        self._check(self._lib.spb200_detect(self._h, _ptr(img), b, c, h, w, capacity, _ptr(count), _ptr(xy), _ptr(conf),
                                            _ptr(desc), _ptr(prob), self._stream()), 'spb200_detect')
        return out

    def set_descriptor_format(self, fmt='fp32'):
        """Element type of the descriptors detect / detect_host return: 'fp32' (the reference's) or 'fp16'."""
        self._check(self._lib.spb200_set_descriptor_format(self._h, {'fp32': 0, 'fp16': 1}[fmt]), 'spb200_set_descriptor_format')
        self.desc_dtype = torch.float16 if fmt == 'fp16' else torch.float32

    desc_dtype = torch.float32

    def alloc_outputs(self, b, capacity, dev, want_desc=True, prob_hw=None):
        count = torch.zeros((b,), dtype=torch.int32, device=dev)
        xy = torch.zeros((b, capacity, 2), dtype=torch.int32, device=dev)
        conf = torch.zeros((b, capacity), dtype=torch.float32, device=dev)
        desc = torch.zeros((b, capacity, 128), dtype=self.desc_dtype, device=dev) if want_desc else None
        prob = torch.empty((b,) + tuple(prob_hw), dtype=torch.float32, device=dev) if prob_hw else None
        return count, xy, conf, desc, prob

    def host_outputs(self, b, capacity, want_desc=True, pinned=False):
        """Host arrays for detect_host / detect_host_wait: (count, xy, conf, desc) numpy, optionally in pinned memory."""
        ddt = np.float16 if self.desc_dtype == torch.float16 else np.float32
        if pinned:
            return (torch.zeros((b,), dtype=torch.int32).pin_memory().numpy(), torch.zeros((b, capacity, 2), dtype=torch.int32).pin_memory().numpy(),
                    torch.zeros((b, capacity), dtype=torch.float32).pin_memory().numpy(),
                    torch.zeros((b, capacity, 128), dtype=self.desc_dtype).pin_memory().numpy() if want_desc else None)
        return (np.zeros((b,), np.int32), np.zeros((b, capacity, 2), np.int32), np.zeros((b, capacity), np.float32),
                np.zeros((b, capacity, 128), ddt) if want_desc else None)

    def detect_host_submit(self, img, capacity, want_desc=True, out=None):
        """First half of detect_host (spb200_detect_host_submit): img float32 B*C*H*W or uint8 B*H*W numpy (host), out =
        (count, xy, conf, desc) numpy arrays the results are downloaded into (host_outputs; allocated here when None).
        Returns a ticket; at most two batches may be in flight.  The arrays are kept alive until detect_host_wait."""
        u8 = img.dtype == np.uint8
        assert img.flags['C_CONTIGUOUS'] and (u8 or img.dtype == np.float32)
        if u8:
            (b, h, w), c = img.shape, 1
        else:
            b, c, h, w = img.shape
        if out is None:
            out = self.host_outputs(b, capacity, want_desc)
        count, xy, conf, desc = out
        ticket = ctypes.c_int()
        self._check(self._lib.spb200_detect_host_submit(self._h, ctypes.c_void_p(img.ctypes.data), int(u8), b, c, h, w, capacity,
                                                        ctypes.c_void_p(count.ctypes.data), ctypes.c_void_p(xy.ctypes.data),
                                                        ctypes.c_void_p(conf.ctypes.data),
                                                        ctypes.c_void_p(desc.ctypes.data if desc is not None else 0),
                                                        ctypes.byref(ticket)), 'spb200_detect_host_submit')
        self._inflight = getattr(self, '_inflight', {})
        self._inflight[ticket.value] = (img, out)
        return ticket.value

    def detect_host_wait(self, ticket):
        """Second half: returns (count, xy, conf, desc) of batch `ticket` once they are complete in the arrays given at submit."""
        img, out = self._inflight.pop(ticket)
        self._check(self._lib.spb200_detect_host_wait(self._h, ticket), 'spb200_detect_host_wait')
        return out

    def detect_host(self, img, capacity, want_desc=True, out=None):
        """img: float32 numpy B*C*H*W (host).  Returns numpy (count, xy, conf, desc)."""
        img = np.ascontiguousarray(img, dtype=np.float32)
        b, c, h, w = img.shape
        if out is None:
            out = self.host_outputs(b, capacity, want_desc)
        count, xy, conf, desc = out
        self._check(self._lib.spb200_detect_host(self._h, ctypes.c_void_p(img.ctypes.data), b, c, h, w, capacity,
                                                 ctypes.c_void_p(count.ctypes.data), ctypes.c_void_p(xy.ctypes.data),
                                                 ctypes.c_void_p(conf.ctypes.data),
                                                 ctypes.c_void_p(desc.ctypes.data if desc is not None else 0)),
                    'spb200_detect_host')
        return out

    def detect_u8(self, img, capacity, want_desc=True, want_prob=False, out=None):
        """img: uint8 CUDA tensor B*H*W (grayscale frames, value k = k/255).  Same outputs as detect()."""
        img = img.contiguous()
        assert img.dtype == torch.uint8 and img.dim() == 3 and img.is_cuda
        b, h, w = img.shape
        if out is None:
            out = self.alloc_outputs(b, capacity, img.device, want_desc, (h, w) if want_prob else None)
        count, xy, conf, desc, prob = out
        self._check(self._lib.spb200_detect_u8(self._h, _ptr(img), b, h, w, capacity, _ptr(count), _ptr(xy), _ptr(conf),
                                               _ptr(desc), _ptr(prob), self._stream()), 'spb200_detect_u8')
        return out

    def detect_host_u8(self, img, capacity, want_desc=True, out=None):
        """img: uint8 numpy B*H*W (host).  Returns numpy (count, xy, conf, desc)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        b, h, w = img.shape
        if out is None:
            out = self.host_outputs(b, capacity, want_desc)
        count, xy, conf, desc = out
        self._check(self._lib.spb200_detect_host_u8(self._h, ctypes.c_void_p(img.ctypes.data), b, h, w, capacity,
                                                    ctypes.c_void_p(count.ctypes.data), ctypes.c_void_p(xy.ctypes.data),
                                                    ctypes.c_void_p(conf.ctypes.data),
                                                    ctypes.c_void_p(desc.ctypes.data if desc is not None else 0)),
                    'spb200_detect_host_u8')
        return out

    def homography_adaptation(self, img, homographies, margin=8, aggregation='sum'):
        """Reference homography_adaptation: img B*C*H*W CUDA fp32, homographies (num, 8) host array.  -> prob B*H*W."""
        img = self._img(img)
        b, c, h, w = img.shape
        hs = np.ascontiguousarray(np.asarray(homographies, dtype=np.float32).reshape(-1, 8))
        prob = torch.empty((b, h, w), dtype=torch.float32, device=img.device)
        self._check(self._lib.spb200_homography_adaptation(self._h, _ptr(img), b, c, h, w, ctypes.c_void_p(hs.ctypes.data),
                                                           hs.shape[0], int(margin), {'sum': 0, 'max': 1}[aggregation],
                                                           _ptr(prob), self._stream()), 'spb200_homography_adaptation')
        return prob

    def match(self, desc_a, count_a, desc_b, count_b, max_dist=0.0):
        """Mutual nearest neighbours (reference get_best_correspondences): desc_* B*cap*D fp32 CUDA tensors, count_* B
        int32.  Returns (match[B,cap] int32: index in b or -1, dist[B,cap] fp32)."""
        desc_a, desc_b = desc_a.contiguous(), desc_b.contiguous()
        b, cap, d = desc_a.shape
        assert desc_b.shape == desc_a.shape
        m = torch.empty((b, cap), dtype=torch.int32, device=desc_a.device)
        dist = torch.empty((b, cap), dtype=torch.float32, device=desc_a.device)
        self._check(self._lib.spb200_match(self._h, _ptr(desc_a), _ptr(count_a), _ptr(desc_b), _ptr(count_b), b, cap, d,
                                           float(max_dist), _ptr(m), _ptr(dist), self._stream()), 'spb200_match')
        return m, dist

    # ---- stage-level ---------------------------------------------------------------------------
    def heatmap_from_logits(self, logits, h, w):
        logits = logits.contiguous()
        b = logits.shape[0]
        prob = torch.empty((b, h, w), dtype=torch.float32, device=logits.device)
        self._check(self._lib.spb200_heatmap_from_logits(self._h, _ptr(logits), b, h, w, _ptr(prob), self._stream()),
                    'spb200_heatmap_from_logits')
        return prob

    def restore_prob_map(self, softmax, h, w):
        """Reference restore_prob_map on an already softmaxed B*65*Hc*Wc CUDA tensor -> B*H*W."""
        softmax = softmax.contiguous()
        b = softmax.shape[0]
        prob = torch.empty((b, h, w), dtype=torch.float32, device=softmax.device)
        self._check(self._lib.spb200_restore_prob_map(self._h, _ptr(softmax), b, h, w, _ptr(prob), self._stream()),
                    'spb200_restore_prob_map')
        return prob

    def preprocess_u8(self, frames, h_out, w_out):
        """The C++ demo's loader on the device: uint8 CUDA frames B*h*w*3 (BGR) or B*h*w (gray) -> uint8 gray B*H*W
        (cv::resize INTER_LINEAR + BGR2GRAY, bit-exact); feed the result to detect_u8."""
        frames = frames.contiguous()
        assert frames.dtype == torch.uint8 and frames.is_cuda and frames.dim() in (3, 4)
        b, h, w = frames.shape[:3]
        c = frames.shape[3] if frames.dim() == 4 else 1
        out = torch.empty((b, h_out, w_out), dtype=torch.uint8, device=frames.device)
        self._check(self._lib.spb200_preprocess_u8(self._h, _ptr(frames), b, h, w, c, _ptr(out), h_out, w_out, self._stream()),
                    'spb200_preprocess_u8')
        return out

    def preprocess_f32(self, frames, h_out, w_out):
        """The Python demo's loader on the device: float32 CUDA frames B*h*w*3 (BGR, [0,1]) -> B*3*H*W RGB
        (make_query_image: ratio-preserving INTER_LINEAR resize + centre crop, then HWC -> CHW)."""
        frames = frames.contiguous()
        assert frames.dtype == torch.float32 and frames.is_cuda and frames.dim() == 4 and frames.shape[3] == 3
        b, h, w = frames.shape[:3]
        out = torch.empty((b, 3, h_out, w_out), dtype=torch.float32, device=frames.device)
        self._check(self._lib.spb200_preprocess_f32(self._h, _ptr(frames), b, h, w, _ptr(out), h_out, w_out, self._stream()),
                    'spb200_preprocess_f32')
        return out

    def nms(self, prob, capacity):
        prob = prob.contiguous()
        b, h, w = prob.shape
        count, xy, conf, _, _ = self.alloc_outputs(b, capacity, prob.device, False)
        self._check(self._lib.spb200_nms(self._h, _ptr(prob), b, h, w, capacity, _ptr(count), _ptr(xy), _ptr(conf),
                                         self._stream()), 'spb200_nms')
        return count, xy, conf

    def sample_descriptors(self, desc_map, h, w, count, xy):
        desc_map = desc_map.contiguous()
        b, d = desc_map.shape[:2]
        cap = xy.shape[1]
        out = torch.zeros((b, cap, d), dtype=torch.float32, device=desc_map.device)
        self._check(self._lib.spb200_sample_descriptors(self._h, _ptr(desc_map), b, d, h, w, cap, _ptr(count), _ptr(xy),
                                                        _ptr(out), self._stream()), 'spb200_sample_descriptors')
        return out

    def export_activation(self, name, batch):
        bid = BUFFERS.index(name)
        c, h, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self._lib.spb200_activation_dims(self._h, bid, ctypes.byref(c), ctypes.byref(h), ctypes.byref(w)),
                    'spb200_activation_dims')
        out = torch.empty((batch, c.value, h.value, w.value), dtype=torch.float32, device='cuda:%d' % self.device)
        self._check(self._lib.spb200_export_activation(self._h, bid, _ptr(out), c.value, self._stream()),
                    'spb200_export_activation')
        return out

    def profile_begin(self):
        self._check(self._lib.spb200_profile_begin(self._h), 'spb200_profile_begin')

    def profile_end(self, max_entries=4096):
        """-> list of (name, ms, flops, bytes) per kernel launch since profile_begin()."""
        names = ctypes.create_string_buffer(64 * max_entries)
        ms = (ctypes.c_float * max_entries)()
        fl = (ctypes.c_double * max_entries)()
        by = (ctypes.c_double * max_entries)()
        n = ctypes.c_int()
        self._check(self._lib.spb200_profile_end(self._h, max_entries, names, ms, fl, by, ctypes.byref(n)),
                    'spb200_profile_end')
        raw = names.raw
        return [(raw[i * 64:(i + 1) * 64].split(b'\0')[0].decode(), float(ms[i]), float(fl[i]), float(by[i]))
                for i in range(n.value)]

    @property
    def kernel_launches(self):
        return self._lib.spb200_kernel_launches(self._h)

    def reset_kernel_launches(self):
        self._lib.spb200_reset_kernel_launches(self._h)
