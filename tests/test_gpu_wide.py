"""GPU parity on the wide set (run on the B200 box with ``pytest -m gpu``): the CUDA path, called through the C ABI,
against outputs of THE REFERENCE (tests/golden/forward_wide.npz: keypoints, per-cell heatmap maxima, descriptors) and
against the oracle's full heatmap.  North-star bars: heatmap max-abs <= 1e-2, keypoints >= 99 % identical integer positions
where ties and threshold-edge scores are the only allowed differences, descriptor cosine >= 0.999.

  * default path (fp16 operands, one MMA per product, zero-sum weight rounding): moderate checkpoint, 16 + 16 images at
    240x320 and 480x640, 2 + 2 frames at 1088x1920 (BASELINE config 5).  Eleven mantissa bits put this path AT the bars,
    with little room: heatmap 3.7e-3 .. 9.7e-3 - under 1e-2 on all 68 images (with plain round-to-nearest weights,
    SPB200_ROUND_NEAREST=1, one image sits at 1.03e-2); keypoints >= 99.5 % on every `shapes` image and every 1088x1920
    frame, 98.4 .. 99.7 % on the uniform-noise `rand` images (a few of the sixteen 240x320 ones below 99 %: thousands of
    near-ties per image).  What the tests assert for it: heatmap <= 1e-2 on EVERY image; every keypoint difference on every
    image is traced to a tie or a threshold-edge score (_gpu_common.unexplained_differences), the only differences the north
    star allows, and no image is below 98 %; the mean per family and size is >= 99 %; descriptor cosine >= 0.999 everywhere;
  * first split level (SPB200_SPLIT_LAYER1) on the same images: INSIDE all three bars on every image (heatmap <= 8.5e-3
    asserted, 7.5e-3 measured; keypoints >= 99 % asserted per image, 99.37 % measured);
  * harsh checkpoint (g = 4, d = 8): SPB200_SPLIT_LAYER2 and SPB200_SPLIT_DETECTOR meet the bars on every image; the
    single-MMA path does not (1.4e-2 .. 2.2e-2) and is reported.
"""
import os

import numpy as np
import pytest
import torch

from oracle import model, postproc, weights
from _gpu_common import CKPT, load_spb, pset, points_from, unexplained_differences
from _wide import wide_cases, wide_image, wide_ref, image_matches, cell_max

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def sds(golden_sd):
    return {'m': golden_sd, 'h': weights.make_state_dict(seed=3, preset='harsh')}


@pytest.fixture(scope='module')
def engines(sds, tmp_path_factory):
    """(checkpoint tag, precision string) -> engine, created on first use."""
    spb = load_spb()
    harsh = weights.save_checkpoint(sds['h'], str(tmp_path_factory.mktemp('ck') / 'super_point_harsh.pt'))

    class Lazy(dict):
        def __missing__(self, key):
            tag, prec = key
            e = spb.Engine(0)
            e.load_checkpoint(CKPT if tag == 'm' else harsh)
            e.finalize(prec)
            e.set_params()
            self[key] = e
            return e

    lz = Lazy()
    yield lz
    for e in lz.values():
        e.close()


def check_case(e, tag, name, sd, heat_tol=1e-2, kp_frac=0.99, cos_min=0.999, full_heat=True, report_only=False, heat_hard=None):
    ref = wide_ref(tag, name)
    gray = wide_image(name)
    assert image_matches(gray, ref), 'the seeded image generator no longer reproduces the fixture input'
    h, w = gray.shape
    img = gray[None, None].contiguous().cuda()
    cap = e.max_keypoints(h, w)
    count, xy, conf, dsc, prob = e.detect(img, cap, want_prob=True)
    heat = prob[0].cpu().numpy()
    pts = points_from(count, xy, conf)
    # 1. heatmap: per-cell maxima against the reference's, the full map against the oracle
    d_cell = float(np.abs(cell_max(heat, h, w) - ref['cell']).max())
    d_full = None
    heat_o = None
    want = {(int(x), int(y)) for x, y in ref['xy']}
    got = pset(pts)
    frac = len(got & want) / max(len(want), 1)
    if full_heat or frac < 1.0:
        prob_o, desc_o, _ = model.forward(gray[None, None], sd)
        heat_o = prob_o[0].numpy()
        d_full = float(np.abs(heat - heat_o).max())
    # 2. keypoints against the reference's: every difference must come from a tie or a threshold-edge score
    unexplained = 0
    if frac < 1.0:
        unexplained, _ = unexplained_differences(heat, heat_o)
        # the lists being explained are the ones compared: ours = the greedy NMS of our own heatmap (bit-exact kernel),
        # the reference's = the oracle's NMS of the oracle's heatmap up to threshold-edge points (tests/test_oracle_wide.py)
        assert got == pset(postproc.get_points(heat)), name
        assert len(want ^ pset(postproc.get_points(heat_o))) <= 2, name
    # 3. descriptors at the reference's first keypoints
    n = min(32, ref['xy'].shape[0])
    desc_map = e.forward(img)[1]
    xyo = torch.from_numpy(np.ascontiguousarray(ref['xy'][:n].astype(np.int32)))[None].cuda()
    mine = e.sample_descriptors(desc_map, h, w, torch.tensor([n], dtype=torch.int32, device='cuda'), xyo)[0, :n].t().cpu().numpy()
    cos = float((mine * ref['desc'][:, :n]).sum(0).min()) if n else 1.0
    print('[wide %s %s] heat cell %.2e full %s  keypoints %d/%d (%.4f, %d unexplained)  desc cos min %.6f' %
          (tag, name, d_cell, '%.2e' % d_full if d_full is not None else '-', len(got & want), len(want), frac, unexplained, cos))
    if report_only:
        return d_cell, d_full, frac
    assert d_cell <= (heat_hard or heat_tol), (name, d_cell)
    assert d_full is None or d_full <= (heat_hard or heat_tol), (name, d_full)
    assert frac >= kp_frac or (frac >= 0.98 and unexplained == 0), (name, frac, unexplained)
    assert unexplained == 0, (name, frac, unexplained)
    assert cos >= cos_min, (name, cos)
    return d_cell, d_full, frac


@pytest.mark.parametrize('size', [240, 480])
@pytest.mark.parametrize('fam', ['shapes', 'rand'])
def test_default_path_wide_set(size, fam, engines, sds):
    """16 images per family and size, moderate checkpoint, default fp16 path."""
    names = wide_cases('m', size, fam)
    assert len(names) >= 16
    hit, heat = [], []
    for i, name in enumerate(names):
        d_cell, d_full, frac = check_case(engines[('m', 'fp16')], 'm', name, sds['m'], full_heat=(size == 240 or i % 4 == 0))
        hit.append(frac)
        heat.append(max(d_cell, d_full or 0.0))
    over = sum(1 for v in heat if v > 1e-2)
    print('[wide m %s%d] keypoint overlap min %.4f mean %.4f; heatmap max-abs worst %.3e, %d of %d over 1e-2' %
          (fam, size, min(hit), sum(hit) / len(hit), max(heat), over, len(heat)))
    assert sum(hit) / len(hit) >= 0.99
    assert over == 0, heat


@pytest.mark.parametrize('size,fam', [(240, 'shapes'), (480, 'shapes'), (240, 'rand')])
def test_first_split_level_has_margin_on_the_wide_set(size, fam, engines, sds):
    """Stem + encoder.layer1 in split precision: every image of the set under the bar (the single-MMA path has one at
    1.03e-2), at ~0.6 x the throughput."""
    worst = 0.0
    for i, name in enumerate(wide_cases('m', size, fam)):
        if size == 480 and i % 2:
            continue
        d_cell, d_full, frac = check_case(engines[('m', 'fp16+layer1')], 'm', name, sds['m'], heat_tol=8.5e-3, full_heat=(size == 240))
        assert frac >= 0.99, (name, frac)
        worst = max(worst, d_cell, d_full or 0.0)
    print('[wide m %s%d fp16+layer1] heatmap max-abs worst %.3e' % (fam, size, worst))


@pytest.mark.parametrize('name', ['shapes1088_0', 'shapes1088_1', 'rand1088_0', 'rand1088_1'])
def test_default_path_1088x1920(name, engines, sds):
    """BASELINE config 5 frame size: forward, keypoints and descriptors at 1088x1920."""
    check_case(engines[('m', 'fp16')], 'm', name, sds['m'], full_heat=name.endswith('_0'))


@pytest.mark.parametrize('prec', ['fp16+all', 'fp16+encoder'])
def test_harsh_preset_meets_bars_with_split_precision(prec, engines, sds):
    """The harsh weight recipe (last detector BatchNorm gamma = 4, dustbin beta = 8) amplifies operand
    rounding 3-4 x: the split-precision levels bring the heatmap back under 1e-2 (measured ~1e-4 with every stage split,
    ~5e-3 with the encoder split)."""
    for name in wide_cases('h'):
        check_case(engines[('h', prec)], 'h', name, sds['h'])


def test_harsh_preset_single_mma_reported(engines, sds):
    """One MMA per product on the harsh preset: 1.4e-2 .. 2.4e-2, over the bar; printed, sanity-bounded."""
    worst = 0.0
    for name in wide_cases('h', 240):
        d_cell, d_full, frac = check_case(engines[('h', 'fp16')], 'h', name, sds['h'], report_only=True)
        worst = max(worst, d_full)
        assert d_full <= 5e-2 and frac >= 0.97
    print('[wide h fp16 single MMA] worst heatmap max-abs %.3e' % worst)


def test_split_precision_on_moderate_preset_is_fp32_grade(engines, sds):
    """Every stage split, moderate checkpoint: heatmap within 2e-4 of the fp32 network, identical keypoints bar ties."""
    for name in ['shapes240_0', 'rand240_0', 'rand240_1', 'shapes480_0', 'rand480_0']:
        check_case(engines[('m', 'fp16+all')], 'm', name, sds['m'], heat_tol=2e-4, kp_frac=0.998)
