"""Oracle: functional fp32 restatement of the reference network (TEST INFRASTRUCTURE).

Works directly on a ``state_dict`` (the 163 tensors the reference's checkpoints hold,
python/src/saveutils.py:54-63) with ``torch.nn.functional`` calls on the CPU; there are no
``nn.Module`` classes here.  Each function names the reference lines it restates.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # nn.BatchNorm2d default, python/src/resnet_blocks.py:8
SOFTMAX_EPS = 0.00001  # python/src/superpoint.py:112
CELL = 8               # python/src/settings.py:7
DESC_DIM = 128         # python/src/superpoint.py:49-50


def _bn(x, sd, name):
    """Eval-mode BatchNorm2d with running statistics."""
    return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'],
                        sd[name + '.weight'], sd[name + '.bias'], False, 0.0, BN_EPS)


def residual_block(x, sd, p, stride):
    """python/src/resnet_blocks.py:14-27.

    relu(bn2(conv1x1(relu(bn1(conv3x3_s(x))))) + shortcut(x)); the shortcut is
    bn(conv1x1_s(x)) when the block owns ``identity_downsample`` (first block of every layer,
    resnet_blocks.py:33-36) and x otherwise.
    """
    y = F.conv2d(x, sd[p + '.conv1.weight'], None, stride, 1)
    y = F.relu(_bn(y, sd, p + '.bn1'))
    y = _bn(F.conv2d(y, sd[p + '.conv2.weight']), sd, p + '.bn2')
    ds = p + '.identity_downsample.0.weight'
    if ds in sd:
        x = _bn(F.conv2d(x, sd[ds], None, stride, 0), sd, p + '.identity_downsample.1')
    return F.relu(y + x)


def layer(x, sd, p, stride):
    """Two-block layer built by make_resnet_layers (resnet_blocks.py:30-40)."""
    x = residual_block(x, sd, p + '.0', stride)
    return residual_block(x, sd, p + '.1', 1)


def encoder(img, sd):
    """python/src/superpoint.py:19-26."""
    x = F.conv2d(img, sd['encoder.conv1.weight'], None, 2, 3)
    x = F.relu(_bn(x, sd, 'encoder.bn1'))
    x = F.max_pool2d(x, 3, 2, 1)
    x = layer(x, sd, 'encoder.layer1', 1)
    return layer(x, sd, 'encoder.layer2', 2)


def detector(feat, sd):
    """python/src/superpoint.py:34-36: two blocks 128->65->65; logits are post-ReLU."""
    return layer(feat, sd, 'detector.layer', 1)


def descriptor(feat, sd):
    """python/src/superpoint.py:52-61."""
    y = layer(feat, sd, 'descriptor.layer_in', 2)
    y = F.conv_transpose2d(y, sd['descriptor.up_sample.weight'], sd['descriptor.up_sample.bias'],
                           stride=2, padding=1, output_padding=1)
    y = F.relu(_bn(y, sd, 'descriptor.bn'))
    y = torch.cat([y, feat], dim=1)
    return layer(y, sd, 'descriptor.layer_out', 1)


def heatmap_from_logits(logits, img_h, img_w):
    """Softmax with epsilon (superpoint.py:111-112) + restore_prob_map (netutils.py:64-75).

    Channel c of cell (i, j) lands on pixel (8 i + c // 8, 8 j + c % 8); the dustbin (c = 64)
    takes part in the sum and is then dropped.
    """
    e = torch.exp(logits)
    p = e / (e.sum(dim=1, keepdim=True) + SOFTMAX_EPS)
    b = p.shape[0]
    hc, wc = int(img_h / CELL), int(img_w / CELL)
    p = p[:, :64].reshape(b, CELL, CELL, hc, wc)      # [b, dy, dx, i, j]
    p = p.permute(0, 3, 1, 4, 2)                      # [b, i, dy, j, dx]
    return p.reshape(b, hc * CELL, wc * CELL)


@torch.no_grad()
def forward(img, sd, descriptor_enabled=True):
    """python/src/superpoint.py:91-115 -> (prob_map B*H*W, desc B*128*Hc*Wc, logits B*65*Hc*Wc).

    ``img`` is B*3*H*W fp32 (a B*1*H*W image is replicated to 3 channels, the reference's
    grayscale convention, python/src/dataset_utils.py:18-20).  With the descriptor head disabled
    (MagicPoint, superpoint.py:105-109) the descriptor map is zeros.
    """
    img = img.float()
    if img.shape[1] == 1:
        img = img.repeat(1, 3, 1, 1)
    h, w = img.shape[-2:]
    feat = encoder(img, sd)
    logits = detector(feat, sd)
    if descriptor_enabled:
        desc = descriptor(feat, sd)
    else:
        desc = torch.zeros((logits.shape[0], DESC_DIM, logits.shape[2], logits.shape[3]))
    return heatmap_from_logits(logits, h, w), desc, logits
