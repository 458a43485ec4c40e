"""Inference settings, same names and defaults as the reference's SuperPointSettings
(python/src/settings.py:3-8) plus the engine knobs the reference does not have."""


class SuperPointSettings:
    def __init__(self):
        self.cuda = True                # the reference defaults to False; this implementation is CUDA only
        self.nms_dist = 4
        self.confidence_thresh = 0.015
        self.nn_thresh = 0.7            # L2 descriptor distance for a good match (unused on this path)
        self.cell = 8
        self.border_remove = 4
        # not in the reference
        self.top_k = 0                  # 0 = every survivor (reference behaviour)
        self.precision = 'fp16'         # 'fp32' (CUDA cores), 'fp16' or 'bf16' (tcgen05)
        self.device = 0

    def read_options(self, opt):
        """python/src/settings.py:33-41 (inference part)."""
        self.cuda = getattr(opt, 'cuda', self.cuda)
        self.nms_dist = opt.nms_dist
        self.confidence_thresh = opt.conf_thresh
        self.nn_thresh = opt.nn_thresh
