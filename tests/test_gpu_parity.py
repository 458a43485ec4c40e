"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

Run on the B200 box with ``pytest -m gpu``.  Bars (BASELINE.json north_star):
  heatmap max-abs <= 1e-2, keypoint sets >= 99 % identical integer positions, descriptor cosine >= 0.999.
The fp32 CUDA-core mode is held to much tighter bounds (it is the same arithmetic as the reference up
to summation order).  Integer work (NMS, sort, top-k) is compared bit-exactly on identical heatmaps.
"""
import os

import numpy as np
import pytest
import torch

from oracle import model, postproc, weights
from _gpu_common import (GOLDEN, CKPT, GOLDEN_CASES, LazyEngines, load_spb, golden_image, pset, points_from,
                         compare_path)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engines():
    e = LazyEngines()
    yield e
    e.close()


# ------------------------------------------------------------------------------------------------
# stage-level kernels against the oracle / golden vectors
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['shapes240_0', 'rand240_1', 'shapes480_0'])
def test_heatmap_kernel(name, engines, golden_sd):
    gray = golden_image(name)
    h, w = gray.shape
    prob, _, logits = model.forward(gray[None, None], golden_sd)
    got = engines['fp32'].heatmap_from_logits(logits.cuda(), h, w).cpu()
    np.testing.assert_allclose(got.numpy(), prob.numpy(), atol=2e-6, rtol=1e-5)
    ref = np.load(os.path.join(GOLDEN, 'forward_%s.npz' % name))
    if 'heatmap' in ref:
        np.testing.assert_allclose(got[0].numpy(), ref['heatmap'], atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize('name', ['shapes240_0', 'shapes240_1', 'shapes240_2', 'rand240_0', 'rand240_1'])
def test_nms_on_golden_heatmaps(name, engines):
    """Same heatmap in -> identical keypoints out as the reference's get_points (ties excepted)."""
    ref = np.load(os.path.join(GOLDEN, 'forward_%s.npz' % name))
    heat = torch.from_numpy(ref['heatmap'])[None].cuda()
    e = engines['fp32']
    count, xy, conf = e.nms(heat, e.max_keypoints(heat.shape[1], heat.shape[2]))
    pts = points_from(count, xy, conf)
    rp = ref['points']
    assert pts.shape == rp.shape
    assert pset(pts) == pset(rp)
    np.testing.assert_array_equal(pts[2].astype(np.float32), rp[2].astype(np.float32))
    # bit-exact against the oracle, order included (same tie rule)
    op = postproc.get_points(ref['heatmap'])
    np.testing.assert_array_equal(pts, op)


def test_nms_known_answers(engines):
    spb = load_spb()
    kat = np.load(os.path.join(GOLDEN, 'nms_kat.npz'))
    tags = sorted({k[:-3] for k in kat.files if k.endswith('_in')})
    e = spb.Engine(0)
    for tag in tags:
        pin, (h, w, dist), pout = kat[tag + '_in'], kat[tag + '_hw'], kat[tag + '_out']
        heat = np.zeros((int(h), int(w)), np.float32)
        if pin.shape[1]:
            heat[pin[1].astype(int), pin[0].astype(int)] = pin[2].astype(np.float32)
        e.set_params(conf_thresh=1e-6, nms_dist=int(dist), border_remove=0)
        count, xy, conf = e.nms(torch.from_numpy(heat)[None].cuda(), max(e.max_keypoints(int(h), int(w), int(dist)), 1))
        got = points_from(count, xy, conf)
        assert pset(got) == pset(pout), tag
        np.testing.assert_array_equal(got, postproc.get_points(heat, 1e-6, int(dist), 0), err_msg=tag)
    e.close()


@pytest.mark.parametrize('shape,batch,dens', [((480, 640), 3, 0.1), ((1088, 1920), 2, 0.05), ((64, 96), 4, 1.0),
                                              ((240, 320), 2, 0.5)])
def test_nms_random_heatmaps_bit_exact(shape, batch, dens, engines):
    h, w = shape
    g = torch.Generator().manual_seed(h + w)
    heat = torch.rand((batch, h, w), generator=g)
    heat = torch.where(torch.rand((batch, h, w), generator=g) < dens, heat, torch.zeros(()))
    heat[0, :, : w // 4] = torch.linspace(0.2, 0.9, w // 4)[None, :]      # ramps: long suppression chains + row ties
    e = engines['fp32']
    for top_k in (0, 300):
        e.set_params(conf_thresh=0.015, nms_dist=4, border_remove=4, top_k=top_k)
        count, xy, conf = e.nms(heat.cuda(), e.max_keypoints(h, w))
        for i in range(batch):
            got = points_from(count, xy, conf, i)
            want = postproc.get_points(heat[i].numpy(), top_k=top_k)
            np.testing.assert_array_equal(got, want)
    e.set_params()


@pytest.mark.parametrize('quant', [0, 64, 4096])
def test_topk_preselection_1080p_bit_exact(quant, engines):
    """BASELINE config 5 shape: tens of thousands of survivors, top_k = 2048.  The finish kernel keeps only the keys
    that can reach the first 2048 places (histogram cut) before sorting; quantised confidences put thousands of
    equal keys into the cut bin (ties by pixel index; with 64 levels the bin overflows and the full sort runs)."""
    h, w = 1088, 1920
    g = torch.Generator().manual_seed(quant + 1)
    heat = torch.rand((1, h, w), generator=g)
    heat = torch.where(torch.rand((1, h, w), generator=g) < 0.3, heat, torch.zeros(()))
    if quant:
        heat = (heat * quant).floor() / quant
    e = engines['fp32']
    for top_k in (2048, 1):
        e.set_params(conf_thresh=0.015, nms_dist=4, border_remove=4, top_k=top_k)
        count, xy, conf = e.nms(heat.cuda(), e.max_keypoints(h, w))
        want_all = postproc.get_points(heat[0].numpy(), top_k=0)
        assert want_all.shape[1] > 8192
        np.testing.assert_array_equal(points_from(count, xy, conf, 0), want_all[:, :top_k])
    e.set_params()


def test_nms_properties_full_size(engines):
    """Size-independent properties at BASELINE config 3 size (64 x 480 x 640)."""
    b, h, w, r = 64, 480, 640, 4
    g = torch.Generator(device='cuda').manual_seed(5)
    heat = torch.rand((b, h, w), generator=g, device='cuda') ** 8
    e = engines['fp32']
    e.set_params(border_remove=0)
    cap = e.max_keypoints(h, w)
    count, xy, conf = e.nms(heat, cap)
    e.set_params()
    kept = torch.zeros((b, h, w), device='cuda')
    for i in range(b):
        n = int(count[i])
        assert n > 0
        c = conf[i, :n]
        assert bool((c[:-1] >= c[1:]).all()), 'sorted by descending confidence'
        x, y = xy[i, :n, 0].long(), xy[i, :n, 1].long()
        assert bool((heat[i, y, x] == c).all())
        kept[i, y, x] = 1
    # no two survivors within the window
    win = torch.nn.functional.avg_pool2d(kept[:, None], 2 * r + 1, 1, r, divisor_override=1)[:, 0]
    assert float((win * kept).max()) == 1.0
    # every candidate is a survivor or has a survivor with >= confidence in its window
    kconf = torch.nn.functional.max_pool2d((kept * heat)[:, None], 2 * r + 1, 1, r)[:, 0]
    cand = heat >= 0.015
    assert bool((kconf[cand] >= heat[cand]).all())


@pytest.mark.parametrize('name', ['shapes240_0', 'rand240_0', 'shapes480_0'])
def test_descriptor_kernel(name, engines, golden_sd):
    ref = np.load(os.path.join(GOLDEN, 'forward_%s.npz' % name))
    gray = golden_image(name)
    h, w = gray.shape
    _, desc, _ = model.forward(gray[None, None], golden_sd)
    rp = ref['points']
    n = ref['descriptors'].shape[1]
    xy = torch.from_numpy(np.ascontiguousarray(rp[:2, :n].T.astype(np.int32)))[None].cuda()
    count = torch.tensor([n], dtype=torch.int32, device='cuda')
    got = engines['fp32'].sample_descriptors(desc.cuda(), h, w, count, xy)[0, :n].t().cpu().numpy()
    np.testing.assert_allclose(got, ref['descriptors'], atol=2e-5)


# ------------------------------------------------------------------------------------------------
# whole path
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_path_fp32_matches_reference(name, engines, golden_sd):
    """fp32 CUDA-core mode against the oracle AND the reference's own stored outputs."""
    gray = golden_image(name)
    compare_path(engines['fp32'], gray, golden_sd, 1e-4, 0.995, 0.99999, 'fp32 ' + name)
    ref = np.load(os.path.join(GOLDEN, 'forward_%s.npz' % name))
    prob, desc, logits = engines['fp32'].forward(gray[None, None].cuda())
    np.testing.assert_allclose(logits[0, :, ::7, ::9].cpu().numpy(), ref['logits_sample'], atol=2e-3, rtol=1e-4)
    np.testing.assert_allclose(desc[0, :, ::7, ::9].cpu().numpy(), ref['desc_sample'], atol=2e-3, rtol=1e-4)
    if 'heatmap' in ref:
        assert float(np.abs(prob[0].cpu().numpy() - ref['heatmap']).max()) <= 1e-4


def test_three_channel_equals_gray(engines):
    gray = golden_image('shapes240_1')[None, None]
    e = engines['fp32']
    p1, d1, l1 = e.forward(gray.cuda())
    p3, d3, l3 = e.forward(gray.repeat(1, 3, 1, 1).cuda())
    assert float((p1 - p3).abs().max()) < 1e-5 and float((d1 - d3).abs().max()) < 1e-3


@pytest.mark.parametrize('prec', ['fp32'])
def test_batch_equals_single_images(prec, engines):
    """Batched semantics == the reference applied per image (inferencewrapper.py:61-66)."""
    e = engines[prec]
    imgs = torch.stack([golden_image('shapes240_%d' % i) for i in range(3)] + [golden_image('rand240_0')])[:, None]
    h, w = imgs.shape[-2:]
    cap = e.max_keypoints(h, w)
    count, xy, conf, dsc, prob = e.detect(imgs.cuda(), cap, want_prob=True)
    count, xy, conf, dsc, prob = [t.clone() for t in (count, xy, conf, dsc, prob)]
    for i in range(imgs.shape[0]):
        c1, xy1, conf1, d1, p1 = e.detect(imgs[i:i + 1].cuda(), cap, want_prob=True)
        assert torch.equal(prob[i], p1[0])
        n = int(c1[0])
        assert int(count[i]) == n
        assert torch.equal(xy[i, :n], xy1[0, :n]) and torch.equal(conf[i, :n], conf1[0, :n])
        assert torch.equal(dsc[i, :n], d1[0, :n])


def test_magicpoint_mode_and_topk(engines, golden_sd):
    e = engines['fp32']
    img = golden_image('shapes240_0')[None, None].cuda()
    cap = e.max_keypoints(240, 320)
    full = e.detect(img, cap)
    full = [t.clone() if t is not None else None for t in full]
    e.set_params(top_k=100)
    c, xy, conf, dsc, _ = e.detect(img, cap)
    assert int(c[0]) == 100 and torch.equal(xy[0, :100], full[1][0, :100]) and torch.equal(conf[0, :100], full[2][0, :100])
    e.set_params(descriptor_enabled=False)          # MagicPoint: detector only (superpoint.py:105-109)
    c2, xy2, conf2, dsc2, _ = e.detect(img, cap)
    assert int(c2[0]) == int(full[0][0]) and torch.equal(xy2[0, :int(c2[0])], full[1][0, :int(c2[0])])
    assert float(dsc2.abs().max()) == 0.0
    prob, desc, logits = e.forward(img)
    assert float(desc.abs().max()) == 0.0
    e.set_params()


def test_empty_image_and_errors(engines):
    e = engines['fp32']
    e.set_params(conf_thresh=0.999)
    c, xy, conf, dsc, _ = e.detect(torch.zeros((1, 1, 64, 96), device='cuda'), e.max_keypoints(64, 96))
    assert int(c[0]) == 0
    e.set_params()
    with pytest.raises(ValueError):
        e.detect(torch.zeros((1, 1, 60, 96), device='cuda'), 10)          # H not a multiple of 16
    with pytest.raises(ValueError):
        e.set_params(nms_dist=99)
    spb = load_spb()
    e2 = spb.Engine(0)
    with pytest.raises(RuntimeError):
        e2.load_checkpoint('/nonexistent/file.pt')
    with pytest.raises(RuntimeError):
        e2.finalize('fp32')                                               # nothing loaded
    e2.close()


def test_dropin_classes(golden_sd):
    """spb200.SuperPoint / InferenceWrapper mirror the reference's classes (superpoint.py:64-115,
    inferencewrapper.py:12-46) and load the checkpoint unchanged."""
    spb = load_spb()
    s = spb.SuperPointSettings()
    s.precision = 'fp32'
    net = spb.SuperPoint(s)
    ck = torch.load(CKPT, map_location='cpu', weights_only=False)
    missing, unexpected = net.load_state_dict(ck['model_state_dict'], strict=True)
    assert not missing and not unexpected
    net.eval()
    gray = golden_image('shapes240_2')
    rgb = gray[None, None].repeat(1, 3, 1, 1)
    prob, desc, logits = net(rgb)
    po, do, lo = model.forward(rgb, golden_sd)
    assert prob.shape == (1, 240, 320) and desc.shape == (1, 128, 30, 40) and logits.shape == (1, 65, 30, 40)
    assert float((prob.cpu() - po).abs().max()) < 1e-4
    one = net(torch.zeros(4, 4))
    assert all(t.shape == (1,) for t in one)
    pts = spb.get_points(prob, 240, 320, s)
    np.testing.assert_array_equal(pts, postproc.get_points(prob.cpu().numpy()))
    dsc = spb.get_descriptors(pts, desc, 240, 320, s)
    np.testing.assert_allclose(dsc, postproc.get_descriptors(pts, desc.cpu().numpy(), 240, 320), atol=2e-5)

    w = spb.InferenceWrapper(CKPT, s)
    img_hwc = np.ascontiguousarray(rgb[0].permute(1, 2, 0).numpy())
    p2, d2 = w.run(img_hwc)
    ref = np.load(os.path.join(GOLDEN, 'forward_shapes240_2.npz'))
    assert p2.shape[0] == 3 and d2.shape[0] == 128 and p2.dtype == np.float64
    inter = len(pset(p2) & pset(ref['points']))
    assert inter >= 0.995 * ref['points'].shape[1]
    with pytest.raises(AssertionError):
        w.run(img_hwc.astype(np.float64))
    with pytest.raises(SystemExit):
        spb.InferenceWrapper('/nonexistent.pt', s)
