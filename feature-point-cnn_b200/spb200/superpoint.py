"""Drop-in for the reference's ``SuperPoint`` module (python/src/superpoint.py:64-115).

Same constructor, same 163 ``state_dict`` keys (so the reference's ``load_checkpoint_for_inference``,
python/src/saveutils.py:6-18, loads into it unchanged), same ``forward`` triple, same
``disable_descriptor`` / ``enable_descriptor`` switches.  The modules below only HOLD parameters;
``forward`` hands them to the sm_100a engine (BatchNorm is folded there) and never runs a PyTorch
convolution.  Inference only: there is no autograd through the engine.
"""
import math

import torch
from torch import nn

from . import ops
from .engine import Engine


class _Conv(nn.Module):
    def __init__(self, cout, cin, k, bias=False, transposed=False):
        super().__init__()
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        self.weight = nn.Parameter(torch.zeros(shape))
        if bias:
            self.bias = nn.Parameter(torch.zeros(cout))
        self.reset_parameters()

    def reset_parameters(self):
        """nn.Conv2d / nn.ConvTranspose2d default initialisation (kaiming_uniform with a = sqrt(5))."""
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if getattr(self, 'bias', None) is not None:
            fan_in = self.weight.shape[1] * self.weight.shape[2] * self.weight.shape[3]
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)


class _BatchNorm(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))

    def reset_parameters(self):
        """nn.BatchNorm2d.reset_parameters."""
        self.running_mean.zero_()
        self.running_var.fill_(1)
        self.num_batches_tracked.zero_()
        nn.init.ones_(self.weight)
        nn.init.zeros_(self.bias)


class _Block(nn.Module):
    """Parameter layout of ResNetBlock (python/src/resnet_blocks.py:5-12)."""

    def __init__(self, cin, cout, downsample):
        super().__init__()
        self.conv1 = _Conv(cout, cin, 3)
        self.bn1 = _BatchNorm(cout)
        self.conv2 = _Conv(cout, cout, 1)
        self.bn2 = _BatchNorm(cout)
        if downsample:
            self.identity_downsample = nn.Sequential(_Conv(cout, cin, 1), _BatchNorm(cout))


def _layer(cin, cout):
    """make_resnet_layers(2, cin, cout, stride) (python/src/resnet_blocks.py:30-40)."""
    return nn.Sequential(_Block(cin, cout, True), _Block(cout, cout, False))


class _Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = _Conv(64, 3, 7)
        self.bn1 = _BatchNorm(64)
        self.layer1 = _layer(64, 64)
        self.layer2 = _layer(64, 128)


class _Detector(nn.Module):
    def __init__(self):
        super().__init__()
        self.layer = _layer(128, 65)


class _Descriptor(nn.Module):
    def __init__(self):
        super().__init__()
        self.layer_in = _layer(128, 256)
        self.up_sample = _Conv(128, 256, 3, bias=True, transposed=True)
        self.bn = _BatchNorm(128)
        self.layer_out = _layer(256, 128)


class SuperPoint(nn.Module):
    def __init__(self, settings):
        super().__init__()
        self.settings = settings
        self.is_descriptor_enabled = True
        self.encoder = _Encoder()
        self.detector = _Detector()
        self.descriptor = _Descriptor()
        self._engine = None
        self._engine_key = None

    def disable_descriptor(self):
        self.is_descriptor_enabled = False

    def enable_descriptor(self):
        self.is_descriptor_enabled = True

    def initialize_descriptor(self):
        """python/src/superpoint.py:86-89: re-initialise the descriptor head's direct children that can be reset (the
        transposed convolution and its BatchNorm; the two residual layers are Sequential containers and are left alone,
        exactly as in the reference)."""
        for layer in self.descriptor.children():
            if hasattr(layer, 'reset_parameters'):
                layer.reset_parameters()

    def _weights_key(self):
        return (getattr(self.settings, 'precision', 'fp16'), tuple(int(p._version) for p in self.state_dict().values()),
                tuple(p.data_ptr() for p in self.state_dict().values()))

    def adopt_engine(self, engine):
        """Use an engine that already holds exactly these parameters (InferenceWrapper.net)."""
        self._engine = engine
        self._engine_key = self._weights_key()

    def engine(self):
        """The engine holding the current parameters (re-uploaded whenever they change)."""
        precision = getattr(self.settings, 'precision', 'fp16')
        key = self._weights_key()
        if self._engine is None:
            self._engine = Engine(getattr(self.settings, 'device', 0))
        if key != self._engine_key:
            self._engine.load_state_dict(self.state_dict())
            self._engine.finalize(precision)
            self._engine_key = key
        s = self.settings
        self._engine.set_params(s.confidence_thresh, s.nms_dist, s.border_remove, getattr(s, 'top_k', 0),
                                self.is_descriptor_enabled)
        return self._engine

    @torch.no_grad()
    def forward(self, image):
        """python/src/superpoint.py:91-115 -> (prob_map B*H*W, desc B*128*Hc*Wc, logits B*65*Hc*Wc), CUDA tensors."""
        if len(image.shape) <= 2:
            return torch.empty((1,)), torch.empty((1,)), torch.empty((1,))
        eng = self.engine()
        image = image.to('cuda:%d' % eng.device, torch.float32)
        return torch.ops.spb200.forward(image, ops.register(eng))
