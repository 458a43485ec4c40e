"""Oracle: seeded synthetic checkpoints and images (TEST INFRASTRUCTURE).

The reference's trained snapshots are not available (reference .MISSING_LARGE_BLOBS:1-2), so
parity and throughput runs use synthetic weights in the exact on-disk format of
``save_checkpoint`` (python/src/saveutils.py:54-63): ``torch.save`` of a dict with ``epoch``,
``model_state_dict`` (163 entries, key names of python/src/superpoint.py:9-17,30-32,40-50 and
python/src/resnet_blocks.py:7-12,33-36), ``optimizer_state_dict`` and ``scaler_state_dict``.

Weight recipe: default PyTorch conv init, every BatchNorm with
gamma~U(0.5,1.5), beta~U(-0.2,0.2), the last detector BN with gamma=g, beta=0, beta[dustbin]=d,
running statistics calibrated by 4 train-mode passes (cumulative average) over uniform-noise
batches.  Presets: moderate (g=1, d=2), harsh (g=4, d=8).  Plain random init is useless here:
the heatmap is ~1/65 everywhere, above the 0.015 threshold.
"""
import math
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

PRESETS = {'moderate': (1.0, 2.0), 'harsh': (4.0, 8.0)}

# (prefix, cin, cout, stride) of the six two-block layers, in state_dict order.
LAYERS = [('encoder.layer1', 64, 64, 1), ('encoder.layer2', 64, 128, 2),
          ('detector.layer', 128, 65, 1),
          ('descriptor.layer_in', 128, 256, 2), ('descriptor.layer_out', 256, 128, 1)]


def _conv_w(gen, cout, cin, k):
    bound = 1.0 / math.sqrt(cin * k * k)      # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
    return (torch.rand((cout, cin, k, k), generator=gen) * 2 - 1) * bound


def _bn_entries(sd, name, c, gen):
    sd[name + '.weight'] = torch.rand(c, generator=gen) + 0.5
    sd[name + '.bias'] = torch.rand(c, generator=gen) * 0.4 - 0.2
    sd[name + '.running_mean'] = torch.zeros(c)
    sd[name + '.running_var'] = torch.ones(c)
    sd[name + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.int64)


def _block_entries(sd, p, cin, cout, first, gen):
    sd[p + '.conv1.weight'] = _conv_w(gen, cout, cin, 3)
    _bn_entries(sd, p + '.bn1', cout, gen)
    sd[p + '.conv2.weight'] = _conv_w(gen, cout, cout, 1)
    _bn_entries(sd, p + '.bn2', cout, gen)
    if first:
        sd[p + '.identity_downsample.0.weight'] = _conv_w(gen, cout, cin, 1)
        _bn_entries(sd, p + '.identity_downsample.1', cout, gen)


def state_dict_keys():
    """The 163 key names, in the reference's registration order."""
    return list(make_state_dict(calibrate=False).keys())


def make_state_dict(seed=0, preset='moderate', calibrate=True):
    g, d = PRESETS[preset]
    gen = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    sd['encoder.conv1.weight'] = _conv_w(gen, 64, 3, 7)
    _bn_entries(sd, 'encoder.bn1', 64, gen)
    for p, cin, cout, _ in LAYERS[:3]:
        _block_entries(sd, p + '.0', cin, cout, True, gen)
        _block_entries(sd, p + '.1', cout, cout, False, gen)
    p, cin, cout, _ = LAYERS[3]
    _block_entries(sd, p + '.0', cin, cout, True, gen)
    _block_entries(sd, p + '.1', cout, cout, False, gen)
    bound = 1.0 / math.sqrt(128 * 9)          # ConvTranspose2d fan_in uses weight.size(1)*k*k
    sd['descriptor.up_sample.weight'] = (torch.rand((256, 128, 3, 3), generator=gen) * 2 - 1) * bound
    sd['descriptor.up_sample.bias'] = (torch.rand(128, generator=gen) * 2 - 1) * bound
    _bn_entries(sd, 'descriptor.bn', 128, gen)
    p, cin, cout, _ = LAYERS[4]
    _block_entries(sd, p + '.0', cin, cout, True, gen)
    _block_entries(sd, p + '.1', cout, cout, False, gen)
    last = 'detector.layer.1.bn2'
    sd[last + '.weight'] = torch.full((65,), float(g))
    sd[last + '.bias'] = torch.zeros(65)
    sd[last + '.bias'][64] = float(d)
    if calibrate:
        _calibrate(sd, gen)
    return sd


def _calibrate(sd, gen, passes=4):
    """Train-mode statistics, cumulative average over `passes` batches of rand(4,3,240,320)."""
    acc = {}

    def bn(x, name):
        mean = x.mean(dim=(0, 2, 3))
        var_b = x.var(dim=(0, 2, 3), unbiased=False)
        var_u = x.var(dim=(0, 2, 3), unbiased=True)
        m, v = acc.setdefault(name, [torch.zeros_like(mean), torch.zeros_like(mean)])
        m += mean / passes
        v += var_u / passes
        xn = (x - mean[None, :, None, None]) / torch.sqrt(var_b[None, :, None, None] + 1e-5)
        return xn * sd[name + '.weight'][None, :, None, None] + sd[name + '.bias'][None, :, None, None]

    def block(x, p, stride):
        y = F.relu(bn(F.conv2d(x, sd[p + '.conv1.weight'], None, stride, 1), p + '.bn1'))
        y = bn(F.conv2d(y, sd[p + '.conv2.weight']), p + '.bn2')
        k = p + '.identity_downsample.0.weight'
        if k in sd:
            x = bn(F.conv2d(x, sd[k], None, stride, 0), p + '.identity_downsample.1')
        return F.relu(y + x)

    def lay(x, p, stride):
        return block(block(x, p + '.0', stride), p + '.1', 1)

    with torch.no_grad():
        for _ in range(passes):
            img = torch.rand((4, 3, 240, 320), generator=gen)
            x = F.relu(bn(F.conv2d(img, sd['encoder.conv1.weight'], None, 2, 3), 'encoder.bn1'))
            x = F.max_pool2d(x, 3, 2, 1)
            feat = lay(lay(x, 'encoder.layer1', 1), 'encoder.layer2', 2)
            lay(feat, 'detector.layer', 1)
            y = lay(feat, 'descriptor.layer_in', 2)
            y = F.conv_transpose2d(y, sd['descriptor.up_sample.weight'], sd['descriptor.up_sample.bias'],
                                   stride=2, padding=1, output_padding=1)
            y = F.relu(bn(y, 'descriptor.bn'))
            lay(torch.cat([y, feat], 1), 'descriptor.layer_out', 1)
    for name, (m, v) in acc.items():
        sd[name + '.running_mean'] = m
        sd[name + '.running_var'] = v
        sd[name + '.num_batches_tracked'] = torch.tensor(passes, dtype=torch.int64)


def save_checkpoint(sd, path, epoch=0):
    """Write `sd` in the save_checkpoint dict format (saveutils.py:57-62)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({'epoch': epoch, 'model_state_dict': sd,
                'optimizer_state_dict': {'state': {}, 'param_groups': []},
                'scaler_state_dict': {}}, path)
    return path


def load_state_dict(path):
    """What load_checkpoint_for_inference reads (saveutils.py:6-18)."""
    ck = torch.load(path, map_location='cpu', weights_only=False)
    return ck['model_state_dict'] if 'model_state_dict' in ck else ck


def _synth():
    """Synthetic input images live with the product (spb200/synth.py) so that bench.py does not need the
    oracle for its inputs; the oracle re-exports them."""
    import importlib
    import sys
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'feature-point-cnn_b200')
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    return importlib.import_module('spb200.synth')


def rand_image(i, h, w):
    return _synth().rand_image(i, h, w)


def shapes_image(i, h, w):
    return _synth().shapes_image(i, h, w)
