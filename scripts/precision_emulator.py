"""Precision emulator: which stages need split-precision operands for the harsh preset.

Emulates the tensor-core path on the CPU: BatchNorm folded, MMA operands rounded to fp16 (optionally as hi + lo
pairs: a.w ~ a_hi.w_hi + a_hi.w_lo + a_lo.w_hi), fp32 accumulation, fp32 bias / shortcut / ReLU, block outputs
stored as fp16 (or as hi + lo).  TEST INFRASTRUCTURE - imports the oracle.
"""
import os
import sys

import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))
from oracle import model, postproc, weights  # noqa: E402


def h(x):
    return x.half().float()


def fold(sd, conv, bn, transposed=False, bias=None):
    s = sd[bn + '.weight'] / torch.sqrt(sd[bn + '.running_var'] + 1e-5)
    w = sd[conv + '.weight']
    w = w * (s[None, :, None, None] if transposed else s[:, None, None, None])
    b = sd[bn + '.bias'] - sd[bn + '.running_mean'] * s
    if bias is not None:
        b = b + sd[bias] * s
    return w, b


def conv(x, w, stride, pad, mode):
    """mode 1: fp16 x fp16; 3: three-term split; 0: fp32.  x is a tuple (hi, lo) or a tensor."""
    if mode == 0:
        xs = x[0] + x[1] if isinstance(x, tuple) else x
        return F.conv2d(xs, w, None, stride, pad)
    if isinstance(x, tuple):
        xh, xl = x
    else:
        xh = h(x); xl = h(x - xh)
    wh = h(w); wl = h(w - wh)
    if mode == 1:
        return F.conv2d(xh, wh, None, stride, pad)
    if mode == 2:      # weights split only
        return F.conv2d(xh, wh, None, stride, pad) + F.conv2d(xh, wl, None, stride, pad)
    return F.conv2d(xh, wh, None, stride, pad) + F.conv2d(xh, wl, None, stride, pad) + F.conv2d(xl, wh, None, stride, pad)


def store(v, mode):
    vh = h(v)
    if mode == 3:
        return (vh, h(v - vh))
    return vh


def full(x):
    return x[0] + x[1] if isinstance(x, tuple) else x


def block(x, sd, p, stride, mode, out_fp32=False, store_mode=None):
    w1, b1 = fold(sd, p + '.conv1', p + '.bn1')
    w2, b2 = fold(sd, p + '.conv2', p + '.bn2')
    y = F.relu(conv(x, w1, stride, 1, mode) + b1[None, :, None, None])
    z = conv(y, w2, 1, 0, mode) + b2[None, :, None, None]
    if p + '.identity_downsample.0.weight' in sd:
        wd, bd = fold(sd, p + '.identity_downsample.0', p + '.identity_downsample.1')
        idn = conv(x, wd, stride, 0, mode) + bd[None, :, None, None]
    else:
        idn = full(x) if mode == 3 else (x[0] if isinstance(x, tuple) else x)
    out = F.relu(z + idn)
    if out_fp32:
        return out
    return store(out, mode if store_mode is None else store_mode)


def emulate(img, sd, modes):
    """modes: dict stage -> 1 / 2 / 3 for 'stem', 'l1', 'l2', 'det'."""
    img3 = img.repeat(1, 3, 1, 1) if img.shape[1] == 1 else img
    w, b = fold(sd, 'encoder.conv1', 'encoder.bn1')
    wg = w.sum(1, keepdim=True)
    g = img3[:, :1]
    ms = modes['stem']
    if ms == 'cur':      # what stem_planes.cu does: image x255 rounded to fp16, weights hi + lo
        gh = h(g * 255.) / 255.
        wh = h(wg); wl = h(wg - wh)
        x = F.conv2d(gh, wh, None, 2, 3) + F.conv2d(gh, wl, None, 2, 3)
    else:
        x = conv(g, wg, 2, 3, ms)
    x = F.max_pool2d(F.relu(x + b[None, :, None, None]), 3, 2, 1)
    x = store(x, modes['l1'])
    x = block(x, sd, 'encoder.layer1.0', 1, modes['l1'])
    x = block(x, sd, 'encoder.layer1.1', 1, modes['l1'], store_mode=modes['l2'])
    x = block(x, sd, 'encoder.layer2.0', 2, modes['l2'])
    feat = block(x, sd, 'encoder.layer2.1', 1, modes['l2'], store_mode=modes['det'])
    d = block(feat, sd, 'detector.layer.0', 1, modes['det'])
    logits = block(d, sd, 'detector.layer.1', 1, modes['det'], out_fp32=True)
    return model.heatmap_from_logits(logits, img.shape[-2], img.shape[-1]), logits


def pset(p):
    return {(int(x), int(y)) for x, y in zip(p[0], p[1])}


def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else 'harsh'
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    sd = weights.make_state_dict(seed=seed, preset=preset)
    imgs = [('shapes0', weights.shapes_image(0, 240, 320)), ('shapes1', weights.shapes_image(1, 240, 320)),
            ('rand0', weights.rand_image(0, 240, 320)), ('rand1', weights.rand_image(1, 240, 320))]
    configs = {
        'all-1': dict(stem='cur', l1=1, l2=1, det=1),
        'all-3 (stem cur)': dict(stem='cur', l1=3, l2=3, det=3),
        'all-3 (stem 3)': dict(stem=3, l1=3, l2=3, det=3),
        'det-3': dict(stem='cur', l1=1, l2=1, det=3),
        'l2+det-3': dict(stem='cur', l1=1, l2=3, det=3),
        'l1+l2-3': dict(stem='cur', l1=3, l2=3, det=1),
        'all-2(w)': dict(stem='cur', l1=2, l2=2, det=2),
    }
    with torch.no_grad():
        for name, gray in imgs:
            img = gray[None, None]
            prob_o, _, logits_o = model.forward(img, sd, descriptor_enabled=False)
            pts_o = pset(postproc.get_points(prob_o.numpy()))
            for cn, modes in configs.items():
                prob, logits = emulate(img, sd, modes)
                pts = pset(postproc.get_points(prob.numpy()))
                print('%-8s %-18s heat %.3e logits %.3e kp %d/%d' % (name, cn, float((prob - prob_o).abs().max()),
                      float((logits - logits_o).abs().max()), len(pts & pts_o), len(pts_o)))
            sys.stdout.flush()


if __name__ == '__main__':
    main()
