"""GPU tests of the split-precision stages (SPB200_SPLIT_*: three MMAs per product, activations and weights as 16-bit
hi + lo pairs): stage outputs against the fp32 CUDA-core path, every input flavour (1 / 3 channels, 8-bit frames),
ragged tile geometry, batches, the host-buffer entry point, bf16 operands.
"""
import numpy as np
import pytest
import torch

from oracle import weights
from _gpu_common import CKPT, load_spb, golden_image

pytestmark = pytest.mark.gpu

STAGES = ['pool', 'l1a', 'l1b', 'l2a', 'feat', 'd0', 'logits']


def make_engine(prec):
    e = load_spb().Engine(0)
    e.load_checkpoint(CKPT)
    e.finalize(prec)
    e.set_params()
    return e


@pytest.fixture(scope='module')
def engines():
    class Lazy(dict):
        def __missing__(self, prec):
            self[prec] = make_engine(prec)
            return self[prec]
    lz = Lazy()
    yield lz
    for e in lz.values():
        e.close()


def stage_errors(e, e32, img, names):
    e32.forward(img)
    e.forward(img)
    out = {}
    for n in names:
        a, b = e32.export_activation(n, img.shape[0]), e.export_activation(n, img.shape[0])
        c = min(a.shape[1], b.shape[1])
        out[n] = float((a[:, :c] - b[:, :c]).abs().max()) / (float(a.abs().max()) + 1e-9)
    return out


@pytest.mark.parametrize('shape', [(240, 320), (208, 272), (112, 400)])
def test_split_levels_stage_by_stage(shape, engines):
    """Each level makes the stages it covers fp32-grade (relative error <= 5e-5 of the stage's largest value, against
    ~1e-3 with one MMA per product) and leaves the others on the single-MMA path; ragged tile geometry included."""
    h, w = shape
    img = torch.stack([weights.rand_image(1, h, w), weights.shapes_image(0, h, w)])[:, None].contiguous().cuda()
    covered = {'fp16+layer1': ['pool', 'l1a'], 'fp16+encoder': ['pool', 'l1a', 'l1b', 'l2a'],
               'fp16+all': ['pool', 'l1a', 'l1b', 'l2a', 'feat', 'd0', 'logits']}
    base = stage_errors(engines['fp16'], engines['fp32'], img, STAGES)
    for prec, exact in covered.items():
        err = stage_errors(engines[prec], engines['fp32'], img, STAGES)
        print('[split %s %dx%d] %s' % (prec, h, w, '  '.join('%s %.1e' % kv for kv in err.items())))
        for n in STAGES:
            if n in exact:
                assert err[n] <= 5e-5, (prec, n, err[n])
            else:
                assert err[n] <= 1.5 * base[n] + 1e-4, (prec, n, err[n], base[n])


def test_split_three_channel_and_8bit_inputs(engines):
    """The split stem takes 1- or 3-channel fp32 images (a gray image replicated to RGB gives the same features as the
    gray-folded kernel up to fp32 rounding) and 8-bit frames (k / 255 as the reference's loaders compute it)."""
    e = engines['fp16+all']
    gray = golden_image('shapes240_1')
    g1 = gray[None, None].contiguous().cuda()
    g3 = g1.repeat(1, 3, 1, 1).contiguous()
    p1 = e.forward(g1)[0].clone()
    p3 = e.forward(g3)[0].clone()
    assert float((p1 - p3).abs().max()) <= 1e-5
    u8 = (gray * 255).round().to(torch.uint8)[None].contiguous().cuda()
    cap = e.max_keypoints(240, 320)
    c8, xy8, conf8, d8, _ = [t.clone() if t is not None else None for t in e.detect_u8(u8, cap)]
    # k / 255 by IEEE division on the host, as the reference's loaders and spb200_detect_u8 do (torch's CUDA division by a
    # scalar multiplies by the reciprocal: one ulp off, which the split stem resolves)
    cf, xyf, conff, df, _ = e.detect((u8.cpu().float() / 255.)[:, None].contiguous().cuda(), cap)
    n = int(cf[0])
    assert int(c8[0]) == n and n > 100
    assert torch.equal(xy8[0, :n], xyf[0, :n]) and torch.equal(d8[0, :n], df[0, :n])


def test_split_batch_equals_single_images_and_host_entry(engines):
    e = engines['fp16+all']
    imgs = torch.stack([golden_image('shapes240_%d' % i) for i in range(3)] + [golden_image('rand240_0')] * 2)[:, None].contiguous()
    cap = e.max_keypoints(240, 320)
    count, xy, conf, dsc, prob = [t.clone() for t in e.detect(imgs.cuda(), cap, want_prob=True)]
    for i in (0, 3, 4):
        c1, xy1, conf1, d1, p1 = e.detect(imgs[i:i + 1].cuda(), cap, want_prob=True)
        n = int(c1[0])
        assert torch.equal(prob[i], p1[0])
        assert int(count[i]) == n and torch.equal(xy[i, :n], xy1[0, :n]) and torch.equal(dsc[i, :n], d1[0, :n])
    torch.cuda.synchronize()
    hc, hxy, hconf, hdsc = e.detect_host(imgs.numpy(), cap)
    np.testing.assert_array_equal(hc, count.cpu().numpy())
    for i in range(imgs.shape[0]):
        n = int(hc[i])
        np.testing.assert_array_equal(hxy[i, :n], xy[i, :n].cpu().numpy())
        np.testing.assert_array_equal(hdsc[i, :n], dsc[i, :n].cpu().numpy())


def test_split_bf16_operands(engines):
    """bf16 hi + lo pairs carry 16 mantissa bits: with every stage split the bf16 path meets the heatmap bar it misses
    with one MMA per product (1.2e-4 by emulation)."""
    img = torch.stack([weights.rand_image(1, 240, 320), weights.shapes_image(0, 240, 320)])[:, None].contiguous().cuda()
    p32 = engines['fp32'].forward(img)[0].clone()
    p = engines['bf16+all'].forward(img)[0]
    d = float((p - p32).abs().max())
    print('[split bf16+all] heatmap max-abs vs fp32 path %.3e' % d)
    assert d <= 2e-3


def test_split_rejects_bad_levels():
    spb = load_spb()
    e = spb.Engine(0)
    e.load_checkpoint(CKPT)
    with pytest.raises(ValueError):
        e.finalize('fp32', split=2)
    with pytest.raises(ValueError):
        e.finalize('fp16', split=7)
    e.finalize('fp16', split='encoder')
    assert e.split_level == 2
    e.close()
