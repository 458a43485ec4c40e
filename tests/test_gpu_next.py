"""GPU tests of the rows next to the hot path: 8-bit frame input (N3) and descriptor matching (N2)."""
import os

import numpy as np
import pytest
import torch

from oracle import matching
from _gpu_common import GOLDEN, LazyEngines, golden_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engines():
    e = LazyEngines()
    yield e
    e.close()


def _frames(names):
    imgs = np.load(os.path.join(GOLDEN, 'images.npz'))
    return np.stack([imgs[n] for n in names]).astype(np.uint8)


@pytest.mark.parametrize('prec', ['fp16', 'fp32'])
def test_u8_frames_equal_float_frames(prec, engines):
    """An 8-bit frame k gives exactly what the reference's loaders feed (k / 255 as fp32): same keypoints, same
    confidences, same descriptors - bit for bit on the tensor-core path (both become the integer k in the 16-bit
    operand), and on the fp32 path (the same division on the device)."""
    u8 = _frames(['s240_0', 's240_1', 's240_2'])
    e = engines[prec]
    cap = e.max_keypoints(240, 320)
    f = torch.from_numpy(u8.astype(np.float32) / 255.)[:, None].cuda()
    c0, xy0, cf0, d0, p0 = [t.clone() for t in e.detect(f, cap, want_prob=True)]
    c1, xy1, cf1, d1, p1 = e.detect_u8(torch.from_numpy(u8).cuda(), cap, want_prob=True)
    assert int(c0.sum()) > 1000
    assert torch.equal(c0, c1) and torch.equal(p0, p1)
    for b in range(3):
        n = int(c0[b])
        assert torch.equal(xy0[b, :n], xy1[b, :n]) and torch.equal(cf0[b, :n], cf1[b, :n]) and torch.equal(d0[b, :n], d1[b, :n])


def test_u8_host_entry_point(engines):
    u8 = _frames(['s240_0', 's240_1'])
    e = engines['fp16']
    cap = e.max_keypoints(240, 320)
    c0, xy0, cf0, d0, _ = e.detect_u8(torch.from_numpy(u8).cuda(), cap)
    c1, xy1, cf1, d1 = e.detect_host_u8(u8, cap)
    assert np.array_equal(c0.cpu().numpy(), c1)
    for b in range(2):
        n = int(c1[b])
        assert np.array_equal(xy0[b, :n].cpu().numpy(), xy1[b, :n]) and np.array_equal(d0[b, :n].cpu().numpy(), d1[b, :n])
    with pytest.raises(Exception):
        e.detect_host_u8(u8[:, :100, :100], cap)            # not a multiple of 16


def _run_match(e, sets_a, sets_b, max_dist=0.0):
    b = len(sets_a)
    cap = max(max(len(s) for s in sets_a), max(len(s) for s in sets_b), 1) + 5
    da = torch.full((b, cap, 128), 7.0)                      # rows beyond the count hold garbage on purpose
    db = torch.full((b, cap, 128), -3.0)
    for i in range(b):
        da[i, :len(sets_a[i])] = torch.from_numpy(sets_a[i])
        db[i, :len(sets_b[i])] = torch.from_numpy(sets_b[i])
    ca = torch.tensor([len(s) for s in sets_a], dtype=torch.int32).cuda()
    cb = torch.tensor([len(s) for s in sets_b], dtype=torch.int32).cuda()
    m, d = e.match(da.cuda(), ca, db.cuda(), cb, max_dist)
    return m.cpu().numpy(), d.cpu().numpy()


def test_match_golden_vectors_of_the_reference(engines):
    """The four golden cases as ONE batch of ragged pairs: identical (queryIdx, trainIdx) to cv2.BFMatcher crossCheck
    as called by the reference, distances within 1e-5."""
    k = np.load(os.path.join(GOLDEN, 'match_kat.npz'))
    names = ['net', 'rand', 'near', 'tiny']
    m, d = _run_match(engines['fp16'], [k[n + '_q'] for n in names], [k[n + '_t'] for n in names])
    for i, n in enumerate(names):
        nq = len(k[n + '_q'])
        qi = np.nonzero(m[i, :nq] >= 0)[0]
        assert list(qi) == list(k[n + '_qidx']), n
        assert list(m[i, qi]) == list(k[n + '_tidx']), n
        np.testing.assert_allclose(d[i, qi], k[n + '_dist'], atol=1e-5)
        assert (m[i, nq:] == -1).all()


@pytest.mark.parametrize('nq,nt,seed', [(1, 1, 0), (64, 64, 1), (65, 129, 2), (700, 3, 3), (5000, 4800, 4)])
def test_match_against_oracle(nq, nt, seed, engines):
    rng = np.random.RandomState(seed)
    a = rng.randn(nq, 128).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    t = rng.randn(nt, 128).astype(np.float32)
    n_copy = min(nq, nt) // 2
    t[:n_copy] = a[rng.permutation(nq)[:n_copy]] + 0.03 * rng.randn(n_copy, 128).astype(np.float32)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    for gate in (0.0, 0.7):                                      # settings.py:6 nn_thresh
        m, d = _run_match(engines['fp16'], [a], [t], gate)
        qi, ti, dist = matching.mutual_nearest(a, t, gate)
        got = np.nonzero(m[0, :nq] >= 0)[0]
        # a pair whose two nearest candidates are within float rounding of each other may legitimately differ
        same = len(set(zip(got, m[0, got])) & set(zip(qi, ti)))
        assert same >= 0.999 * len(qi) and abs(len(got) - len(qi)) <= max(1, len(qi) // 1000)
        common = np.intersect1d(got, qi)
        np.testing.assert_allclose(d[0, common], dist[np.searchsorted(qi, common)], atol=2e-5)


def test_match_256_wide_descriptors_cuda_core_kernel(engines):
    """D = 256 (the stale C++ demo's descriptor width, cpp/src/torchutis.h:11) takes the fp32 CUDA-core kernel."""
    rng = np.random.RandomState(9)
    a = rng.randn(300, 256).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    t = np.vstack([a[rng.permutation(300)[:120]] + 0.02 * rng.randn(120, 256).astype(np.float32), rng.randn(137, 256).astype(np.float32)])
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    e = engines['fp16']
    cap = 320
    da, db = torch.zeros((1, cap, 256)), torch.zeros((1, cap, 256))
    da[0, :300], db[0, :257] = torch.from_numpy(a), torch.from_numpy(t)
    m, d = e.match(da.cuda(), torch.tensor([300], dtype=torch.int32).cuda(), db.cuda(), torch.tensor([257], dtype=torch.int32).cuda(), 0.7)
    qi, ti, dist = matching.mutual_nearest(a, t, 0.7)
    got = np.nonzero(m[0, :300].cpu().numpy() >= 0)[0]
    assert list(got) == list(qi) and list(m[0, got].cpu().numpy()) == list(ti)
    np.testing.assert_allclose(d[0, got].cpu().numpy(), dist, atol=2e-5)


def test_match_empty_sets_and_linear_pipeline(engines):
    """count = 0 on either side gives no matches; detect -> match on two frames of the same scene runs end to end
    and every match is mutual."""
    e = engines['fp16']
    a = np.eye(8, 128, dtype=np.float32)
    m, _ = _run_match(e, [a[:0], a], [a, a[:0]])
    assert (m == -1).all()
    u8 = _frames(['s240_0', 's240_0'])
    u8[1] = np.roll(u8[1], 16, axis=1)                         # the same scene shifted by two cells (the network is
                                                               # equivariant to shifts by its total stride)
    cap = e.max_keypoints(240, 320)
    count, xy, conf, desc, _ = e.detect_u8(torch.from_numpy(u8).cuda(), cap)
    m, d = e.match(desc[0:1], count[0:1], desc[1:2], count[1:2], 0.7)
    mb, _ = e.match(desc[1:2], count[1:2], desc[0:1], count[0:1], 0.7)
    m, mb = m[0].cpu().numpy(), mb[0].cpu().numpy()
    qi = np.nonzero(m >= 0)[0]
    assert len(qi) > 100
    assert (mb[m[qi]] == qi).all()
    dx = (xy[1, m[qi], 0] - xy[0, qi, 0]).cpu().numpy()
    assert np.median(dx) == 16


# ---- N1: homography adaptation --------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['default', 'preprocess'])
def test_homography_adaptation_matches_reference_golden(name, engines):
    """spb200_homography_adaptation against the map the REFERENCE's homography_adaptation produced for the same
    images and the same sampled homographies: fp32 engine within 1e-3 everywhere (bilinear weights and nearest
    rounding at validity-mask edges are the only differences), fp16 tensor-core engine within the heatmap bar 1e-2."""
    k = np.load(os.path.join(GOLDEN, 'homography_kat.npz'))
    imgs = np.load(os.path.join(GOLDEN, 'images.npz'))
    x = torch.from_numpy(np.stack([imgs[str(i)] for i in k[name + '_ids']]).astype(np.float32) / 255.)[:, None].cuda()
    num, margin, agg = [int(v) for v in k[name + '_cfg']]
    ref = k[name + '_prob']
    for prec, tol in (('fp32', 1e-3), ('fp16', 1e-2)):
        p = engines[prec].homography_adaptation(x, k[name + '_H'], margin, 'sum' if agg == 0 else 'max').cpu().numpy()
        d = np.abs(p - ref)
        print('[homography adaptation %s %s] max abs %.3e, %d px above 1e-4' % (name, prec, float(d.max()), int((d > 1e-4).sum())))
        assert float(d.max()) <= tol, (name, prec)
        assert ((p == 0) == (ref == 0)).mean() > 0.9995          # the count >= num // 3 gate


def test_homography_adaptation_against_oracle_and_wrapper(engines, tmp_path):
    """Other sizes / counts against the oracle (incl. num = 0, margin = 0, 'max'), and the drop-in wrapper entry
    point run_with_homography_adaptation end to end."""
    from oracle import homography as oh, model, weights
    from _gpu_common import CKPT, load_spb
    spb = load_spb()
    from spb200 import homographies as hg
    sd = weights.load_state_dict(CKPT)
    rng = np.random.default_rng(5)
    e = engines['fp32']
    for (b, h, w, num, margin, agg) in [(1, 96, 128, 0, 8, 'sum'), (2, 96, 128, 4, 0, 'sum'), (1, 128, 96, 5, 3, 'max')]:
        x = torch.rand((b, 1, h, w), generator=torch.Generator().manual_seed(h + num))
        cfg = hg.HomographyConfig()
        cfg.num = num
        hs = hg.sample_homographies((h, w), cfg, rng)
        want = oh.homography_adaptation(x, lambda im: model.forward(im, sd)[0], hs, margin, agg).numpy()
        got = e.homography_adaptation(x.cuda(), hs, margin, agg).cpu().numpy()
        assert float(np.abs(got - want).max()) <= 1e-3, (b, h, w, num, margin, agg)
    settings = spb.SuperPointSettings()
    settings.precision = 'fp32'
    wrap = spb.InferenceWrapper(CKPT, settings)
    cfg = hg.HomographyConfig()
    cfg.num = 6
    img = np.repeat(golden_image('shapes240_0').numpy()[:, :, None], 3, axis=2).astype(np.float32)
    pts = wrap.run_with_homography_adaptation(img, cfg, rng=rng)
    assert len(pts) == 1 and pts[0].shape[0] == 3 and pts[0].shape[1] > 100
    assert (np.diff(pts[0][2]) <= 0).all()                        # descending confidence
    wrap.engine.close()


def test_cpp_facade_demo_runs_float_and_8bit_frames():
    """The C++ mirror of the reference's wrapper (cpp/superpoint.h: superpoint::SuperPoint, ProcessFrame, ProcessFrame8)
    through its headless demo binary: loads the checkpoint, finds keypoints, and the 8-bit frame gives the same ones."""
    import subprocess
    from _gpu_common import CKPT
    exe = os.path.join(os.path.dirname(GOLDEN), '..', 'feature-point-cnn_b200', 'cpp', 'demo')
    r = subprocess.run([os.path.abspath(exe), CKPT, '240', '320'], capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0
    assert 'descriptor dim 128' in r.stdout and 'identical to the float frame' in r.stdout


def test_engine_loads_params_file_of_the_reference_export(tmp_path):
    """N4 (weights half): the '<name>_params.pt' InferenceWrapper.trace writes (python/src/inferencewrapper.py:89-91,
    keys without their module prefix) gives the same network as the full checkpoint."""
    from _gpu_common import CKPT, load_spb
    spb = load_spb()
    sd = torch.load(CKPT, map_location='cpu', weights_only=False)['model_state_dict']
    p = str(tmp_path / 'super_point_params.pt')
    torch.save({('.'.join(k.split('.')[1:])): v for k, v in sd.items()}, p, _use_new_zipfile_serialization=True)
    img = golden_image('shapes240_0')[None, None].cuda()
    outs = []
    for path in (CKPT, p):
        e = spb.Engine(0)
        e.load_checkpoint(path)
        e.finalize('fp16')
        e.set_params()
        outs.append([t.clone() for t in e.forward(img)])
        e.close()
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_frame_loaders_match_opencv_and_the_reference(engines):
    """N3: the demos' loaders on the device against tests/golden/preproc_kat.npz - cv::resize + BGR2GRAY on 8-bit frames
    (cpp/src/camera.cc:12-23) bit for bit; the reference's make_query_image (python/src/inference.py:72-85) + HWC -> CHW to
    float rounding (OpenCV's float upscale differs from the two-pass form in the last bits)."""
    import os
    import numpy as np
    kat = np.load(os.path.join(GOLDEN, 'preproc_kat.npz'))
    e = engines['fp16']
    n8 = len([k for k in kat.files if k.startswith('u8_') and k.endswith('_in')])
    nf = len([k for k in kat.files if k.startswith('f32_') and k.endswith('_in')])
    assert n8 >= 5 and nf >= 5
    for k in range(n8):
        f, want = kat['u8_%d_in' % k], kat['u8_%d_out' % k]
        got = e.preprocess_u8(torch.from_numpy(np.stack([f, f[::-1].copy()])).cuda(), want.shape[0], want.shape[1]).cpu().numpy()
        np.testing.assert_array_equal(got[0], want, err_msg='u8 case %d' % k)
        assert got[1].shape == want.shape
    for k in range(nf):
        f, want = kat['f32_%d_in' % k], kat['f32_%d_out' % k]
        got = e.preprocess_f32(torch.from_numpy(f[None]).cuda(), want.shape[1], want.shape[2]).cpu().numpy()[0]
        assert float(np.abs(got - want).max()) <= 1e-5, ('f32 case %d' % k, float(np.abs(got - want).max()))
    # the loaded frames feed the path: 8-bit gray frames -> detect_u8
    f = kat['u8_1_in']
    gray = e.preprocess_u8(torch.from_numpy(f[None]).cuda(), 112, 160)
    count, xy, conf, desc, _ = e.detect_u8(gray, e.max_keypoints(112, 160))
    assert int(count[0]) >= 0


def test_headless_cli_and_pseudo_label_files(tmp_path):
    """N3: `python -m spb200.main inference` over a folder of frames (python/main.py:9-99, python/src/inference.py:10-69
    without camera and window) writes per frame what InferenceWrapper.run returns plus the mutual-nearest-neighbour matches
    to the previous frame; the pseudo-labelling writer produces the reference's .npz entries (preprocess_coco.py:74:
    image (3, H, W) float32, points (3, N)) that dataset_utils.read_dataset_item consumes."""
    import cv2
    import spb200
    from spb200 import main as cli, preprocess_coco as pc, homographies as hg
    from _gpu_common import CKPT
    imgs = np.load(os.path.join(GOLDEN, 'images.npz'))
    frames = tmp_path / 'frames'
    frames.mkdir()
    for i, key in enumerate(['s240_0', 's240_1', 's240_0']):
        cv2.imwrite(str(frames / ('f%d.png' % i)), cv2.cvtColor(imgs[key], cv2.COLOR_GRAY2BGR))
    out = tmp_path / 'features'
    assert cli.main(['--H', '240', '--W', '320', 'inference', '--weights-path', CKPT, '--images', str(frames), '--out', str(out),
                     '--out-file-name', str(tmp_path / 'exported')]) == 0
    files = sorted(os.listdir(out))
    assert files == ['f0.npz', 'f1.npz', 'f2.npz']
    s = spb200.SuperPointSettings()
    w = spb200.InferenceWrapper(CKPT, s)
    f0 = np.load(out / 'f0.npz')
    rgb = np.repeat(imgs['s240_0'].astype(np.float32)[:, :, None] / 255., 3, axis=2)          # same size: the loader is the identity
    pts, dsc = w.run(rgb)
    assert f0['points'].shape == pts.shape and pts.shape[1] > 500
    np.testing.assert_array_equal(f0['points'], pts)
    np.testing.assert_allclose(f0['descriptors'], dsc, atol=1e-6)
    f2 = np.load(out / 'f2.npz')
    m = f2['matches']
    assert m.shape[1] == 2 and len(m) > 10 and m[:, 0].max() < f2["points"].shape[1]
    # frames 0 and 2 are the same image: matching frame 2 against frame 1 equals matching frame 0 against frame 1
    f1 = np.load(out / 'f1.npz')
    qi, ti, _ = matching.mutual_nearest(f2['descriptors'].T, f1['descriptors'].T, 0.7)        # settings.nn_thresh
    np.testing.assert_array_equal(m[:, 0], qi)
    np.testing.assert_array_equal(m[:, 1], ti)
    # the exported weights file loads back (InferenceWrapper.trace, python/src/inferencewrapper.py:89-91)
    w2 = spb200.InferenceWrapper(str(tmp_path / 'exported_params.pt'), s)
    p2, _ = w2.run(rgb)
    np.testing.assert_array_equal(p2, pts)
    # pseudo labels
    cfg = hg.HomographyConfig()
    cfg.init_for_preprocess()
    cfg.num = 5
    w.descriptor_enabled = False
    written = pc.preprocess_coco_folder(sorted(str(p) for p in frames.glob('*.png'))[:2], w, cfg, tmp_path / 'labels', (240, 320),
                                        rng=np.random.default_rng(1))
    assert len(written) == 2
    item = np.load(written[0])
    assert sorted(item.files) == ['image', 'points']
    assert item['image'].shape == (3, 240, 320) and item['image'].dtype == np.float32 and 0 <= item['image'].min() and item['image'].max() <= 1
    assert item['points'].shape[0] == 3 and item['points'].shape[1] > 50
    assert np.all(np.diff(item['points'][2]) <= 0)                 # descending confidence, as get_points returns
