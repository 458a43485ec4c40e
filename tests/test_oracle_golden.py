"""Pin the oracle to outputs of the reference itself (tests/golden, made by make_golden.py).

CPU only.  Every oracle function is checked against what the reference's own code produced for
the same inputs: SuperPoint.forward (python/src/superpoint.py:91-115), get_points / corners_nms
(python/src/netutils.py:78-100, python/src/nms.py:4-53), get_descriptors (netutils.py:103-121).
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import model, postproc, weights


def _image_for(name, golden_dir):
    imgs = np.load(os.path.join(golden_dir, 'images.npz'))
    fam, idx = name.rsplit('_', 1)
    if fam.startswith('shapes'):
        key = 's%s_%s' % (fam[len('shapes'):], idx)
        return torch.from_numpy(imgs[key].astype(np.float32) / 255.)
    h = int(fam[len('rand'):])
    return weights.rand_image(int(idx), h, h * 4 // 3)


def _cases(golden_dir):
    return sorted(os.path.basename(p)[len('forward_'):-4]
                  for p in glob.glob(os.path.join(golden_dir, 'forward_*.npz')))


def _point_set(pts):
    return {(int(x), int(y)) for x, y in zip(pts[0], pts[1])}


def test_state_dict_has_reference_keys(golden_sd):
    assert len(golden_sd) == 163
    assert list(golden_sd.keys()) == weights.state_dict_keys()


@pytest.mark.parametrize('name', ['shapes240_0', 'shapes240_1', 'shapes240_2', 'rand240_0', 'rand240_1',
                                  'shapes480_0'])
def test_forward_and_postproc_match_reference(name, golden_dir, golden_sd):
    ref = np.load(os.path.join(golden_dir, 'forward_%s.npz' % name))
    gray = _image_for(name, golden_dir)
    h, w = gray.shape
    prob, desc, logits = model.forward(gray[None, None], golden_sd)
    # forward: same ATen ops in the same order -> tiny differences only
    np.testing.assert_allclose(logits[0, :, ::7, ::9].numpy(), ref['logits_sample'], atol=2e-4, rtol=1e-5)
    np.testing.assert_allclose(desc[0, :, ::7, ::9].numpy(), ref['desc_sample'], atol=2e-4, rtol=1e-5)
    assert abs(logits.double().sum().item() - float(ref['logits_sum'])) <= 1e-6 * abs(float(ref['logits_sum'])) + 1e-2
    if 'heatmap' in ref:
        np.testing.assert_allclose(prob[0].numpy(), ref['heatmap'], atol=1e-6)
        heat = ref['heatmap']
    else:
        np.testing.assert_allclose(prob[0, ::3, ::3].numpy(), ref['heatmap_sample'], atol=1e-6)
        heat = prob[0].numpy()
    # keypoints from the REFERENCE heatmap (when stored) so that NMS is compared like for like
    pts = postproc.get_points(heat)
    rp = ref['points']
    assert pts.shape == rp.shape
    assert _point_set(pts) == _point_set(rp)
    np.testing.assert_array_equal(pts[2], rp[2])            # same descending confidences
    distinct = np.concatenate([[True], np.diff(rp[2]) != 0]) & np.concatenate([np.diff(rp[2]) != 0, [True]])
    np.testing.assert_array_equal(pts[:2, distinct], rp[:2, distinct])   # order differs on ties only
    # descriptors at the reference's points
    dsc = postproc.get_descriptors(rp, desc.numpy(), h, w)
    n = ref['descriptors'].shape[1]
    np.testing.assert_allclose(dsc[:, :n], ref['descriptors'], atol=2e-5)
    assert np.allclose(np.linalg.norm(dsc, axis=0), 1.0, atol=1e-5)


def test_nms_known_answers(golden_dir):
    kat = np.load(os.path.join(golden_dir, 'nms_kat.npz'))
    tags = sorted({k[:-3] for k in kat.files if k.endswith('_in')})
    assert len(tags) >= 9
    for tag in tags:
        pin, (h, w, dist), pout = kat[tag + '_in'], kat[tag + '_hw'], kat[tag + '_out']
        heat = np.zeros((h, w), np.float32)
        if pin.shape[1]:
            heat[pin[1].astype(int), pin[0].astype(int)] = pin[2].astype(np.float32)
        got = postproc.get_points(heat, conf_thresh=1e-6, nms_dist=int(dist), border_remove=0)
        assert _point_set(got) == _point_set(pout), tag
        if pout.shape[1] > 1:
            np.testing.assert_allclose(got[2], pout[2].astype(np.float32), rtol=0, atol=0, err_msg=tag)
        py = postproc.greedy_nms_python(pin, int(h), int(w), int(dist))
        assert _point_set(py) == _point_set(pout), tag


def test_label_round_trip(golden_dir):
    """make_points_labels -> make_prob_map_from_labels -> get_points (python/tests/synthetic-test.py:27-28)."""
    kat = np.load(os.path.join(golden_dir, 'label_kat.npz'))
    got = postproc.get_points(kat['prob_map'])
    assert _point_set(got) == _point_set(kat['points'])
    labelled = {(int(x), int(y)) for y, x in kat['label_points']}
    assert _point_set(got) <= labelled


def test_restore_prob_map_mapping():
    """Channel c of cell (i,j) -> pixel (8i + c//8, 8j + c%8) (netutils.py:64-75)."""
    hc, wc = 3, 5
    logits = torch.zeros(1, 65, hc, wc)
    logits[0, 19, 1, 2] = 30.0
    prob = model.heatmap_from_logits(logits, hc * 8, wc * 8)
    y, x = np.unravel_index(int(prob[0].argmax()), (hc * 8, wc * 8))
    assert (y, x) == (8 * 1 + 19 // 8, 8 * 2 + 19 % 8)


def test_empty_and_single():
    heat = np.zeros((16, 16), np.float32)
    assert postproc.get_points(heat).shape == (3, 0)
    assert postproc.get_descriptors(np.zeros((3, 0)), np.ones((1, 128, 2, 2), np.float32), 16, 16).shape == (128, 0)
    heat[8, 9] = 0.5
    pts = postproc.get_points(heat)
    assert pts.shape == (3, 1) and (pts[0, 0], pts[1, 0]) == (9, 8)
    heat[2, 9] = 0.9     # in the border band: suppresses nothing here (distance 6) and is removed
    assert postproc.get_points(heat).shape == (3, 1)
    heat[5, 9] = 0.4     # dy = 3 from the border point: suppressed by it although that one is removed
    assert _point_set(postproc.get_points(heat)) == {(9, 8)}


def test_synthetic_recipe_is_peaky():
    sd = weights.make_state_dict(seed=1, preset='moderate')
    assert list(sd.keys()) == weights.state_dict_keys()
    img = weights.shapes_image(0, 240, 320)
    prob, desc, logits = model.forward(img[None, None], sd)
    frac = float((prob >= 0.015).float().mean())
    assert 0.005 < frac < 0.4, frac
    assert desc.shape == (1, 128, 30, 40) and logits.shape == (1, 65, 30, 40)
