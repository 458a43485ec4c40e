"""CPU-only checks of the C-ABI library and the host logic (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, 'tests', 'golden')


@pytest.fixture(scope='module')
def lib():
    import sys
    sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))
    import build as spb_build       # feature-point-cnn_b200/build.py
    spb_build.build()
    from spb200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from spb200 import _lib
    header = open(os.path.join(REPO, 'include', 'spb200.h')).read()
    declared = set(re.findall(r'SPB200_API[^;(]*?\b(spb200_\w+)\s*\(', header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert len(declared) >= 22


def test_host_only_entry_points(lib):
    assert lib.spb200_max_keypoints(480, 640, 4) == 96 * 128
    assert lib.spb200_max_keypoints(1088, 1920, 4) == 218 * 384
    assert lib.spb200_max_keypoints(16, 16, 0) == 256
    assert lib.spb200_descriptor_dim(None) == 128
    assert lib.spb200_kernel_launches(None) == 0


def test_create_fails_loudly_without_gpu(lib):
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    h = ctypes.c_void_p()
    rc = lib.spb200_create(0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert lib.spb200_last_error(None)
    import spb200
    with pytest.raises(spb200.Spb200Error):
        spb200.Engine(0)


def test_checkpoint_reader_matches_torch_load(lib):
    """The library's own ZIP + pickle reader against torch.load on the reference-written checkpoint."""
    path = os.path.join(GOLDEN, 'super_point.pt').encode()
    sd = torch.load(path.decode(), map_location='cpu', weights_only=False)['model_state_dict']
    assert lib.spb200_checkpoint_num_tensors(path) == len(sd) == 163
    shape = (ctypes.c_int64 * 8)()
    rank = ctypes.c_int()
    for key, t in sd.items():
        buf = np.empty(max(t.numel(), 1), np.float32)
        rc = lib.spb200_checkpoint_tensor(path, key.encode(), ctypes.c_void_p(buf.ctypes.data), buf.size, shape,
                                          ctypes.byref(rank))
        assert rc == 0, key
        assert tuple(shape[:rank.value]) == tuple(t.shape), key
        np.testing.assert_array_equal(buf[:t.numel()], t.to(torch.float32).reshape(-1).numpy(), err_msg=key)
        if key.endswith('conv1.weight') and 'layer2.1' in key:
            break   # a representative prefix is enough for every dtype/shape kind; keep the test fast
    assert lib.spb200_checkpoint_tensor(path, b'no.such.key', None, 0, shape, ctypes.byref(rank)) != 0
    assert lib.spb200_checkpoint_num_tensors(b'/nonexistent.pt') < 0


def test_checkpoint_reader_variants(lib, tmp_path):
    """Bare state_dict (inferencewrapper.py:89-91), non-contiguous / half / double tensors, garbage files."""
    t = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    sd = {'a.weight': t.permute(2, 0, 1), 'b.half': t.half()[1:], 'c.double': t.double() * 0.5,
          'd.long': torch.tensor(7, dtype=torch.int64), 'e.bf16': t.bfloat16()}
    p = str(tmp_path / 'bare.pt')
    torch.save(sd, p)
    assert lib.spb200_checkpoint_num_tensors(p.encode()) == 5
    shape = (ctypes.c_int64 * 8)()
    rank = ctypes.c_int()
    for key, ten in sd.items():
        buf = np.empty(max(ten.numel(), 1), np.float32)
        assert lib.spb200_checkpoint_tensor(p.encode(), key.encode(), ctypes.c_void_p(buf.ctypes.data), buf.size, shape,
                                            ctypes.byref(rank)) == 0
        assert tuple(shape[:rank.value]) == tuple(ten.shape)
        np.testing.assert_array_equal(buf[:ten.numel()], ten.to(torch.float32).contiguous().reshape(-1).numpy())
    bad = tmp_path / 'garbage.pt'
    bad.write_bytes(b'not a zip file at all' * 10)
    assert lib.spb200_checkpoint_num_tensors(str(bad).encode()) < 0
    q = str(tmp_path / 'legacy.pt')
    torch.save(sd, q, _use_new_zipfile_serialization=False)
    assert lib.spb200_checkpoint_num_tensors(q.encode()) < 0     # documented: legacy (non-zip) format unsupported


def test_params_file_of_inferencewrapper_trace(lib, tmp_path):
    """InferenceWrapper.trace (python/src/inferencewrapper.py:83-91) saves '<name>_params.pt' with the first component of
    every key stripped; the loader puts the module prefixes back (the weights half of it)."""
    sd = torch.load(os.path.join(GOLDEN, 'super_point.pt'), map_location='cpu', weights_only=False)['model_state_dict']
    stripped = {('.'.join(k.split('.')[1:])): v for k, v in sd.items()}           # the reference's own expression
    assert len(stripped) == len(sd)                                                # still unambiguous
    p = str(tmp_path / 'super_point_params.pt')
    torch.save(stripped, p, _use_new_zipfile_serialization=True)
    assert lib.spb200_checkpoint_num_tensors(p.encode()) == 163
    shape = (ctypes.c_int64 * 8)()
    rank = ctypes.c_int()
    for key in ('encoder.conv1.weight', 'encoder.bn1.running_var', 'detector.layer.1.bn2.bias', 'descriptor.up_sample.bias',
                'descriptor.bn.weight', 'descriptor.layer_out.0.identity_downsample.0.weight'):
        t = sd[key]
        buf = np.empty(t.numel(), np.float32)
        assert lib.spb200_checkpoint_tensor(p.encode(), key.encode(), ctypes.c_void_p(buf.ctypes.data), buf.size, shape,
                                            ctypes.byref(rank)) == 0, key
        np.testing.assert_array_equal(buf, t.to(torch.float32).reshape(-1).numpy(), err_msg=key)


def test_dropin_module_has_reference_state_dict_keys():
    import spb200
    from oracle import weights
    net = spb200.SuperPoint(spb200.SuperPointSettings())
    assert list(net.state_dict().keys()) == weights.state_dict_keys()
    sd = weights.load_state_dict(os.path.join(GOLDEN, 'super_point.pt'))
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    for k, v in net.state_dict().items():
        assert v.shape == sd[k].shape, k
    s = spb200.SuperPointSettings()
    assert (s.nms_dist, s.confidence_thresh, s.cell, s.border_remove, s.nn_thresh) == (4, 0.015, 8, 4, 0.7)
