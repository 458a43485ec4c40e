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


def test_checkpoint_reader_survives_truncated_and_corrupted_archives(lib, tmp_path):
    """A truncated or damaged checkpoint must come back as an error code with a message, never as an out-of-bounds read:
    every prefix class of a small archive, single-byte corruptions of its pickle, central directory and end record, and
    hostile 64-bit sizes (the reader trusts nothing it reads from the file)."""
    t = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    good = tmp_path / 'small.pt'
    torch.save({'model_state_dict': {'a.weight': t, 'b.bias': t[0, 0], 'c.half': t.half()}}, str(good))
    raw = good.read_bytes()
    assert lib.spb200_checkpoint_num_tensors(str(good).encode()) == 3
    bad = tmp_path / 'bad.pt'
    for cut in list(range(0, 64)) + list(range(64, len(raw), 37)) + [len(raw) - k for k in range(1, 60)]:
        bad.write_bytes(raw[:cut])
        assert lib.spb200_checkpoint_num_tensors(str(bad).encode()) < 0, cut
        assert lib.spb200_last_error(None)
    rs = np.random.RandomState(0)
    eocd = raw.rfind(b'PK\x05\x06')
    cdir = raw.find(b'PK\x01\x02')
    pkl = raw.find(b'data.pkl') + 8
    spots = (list(range(eocd, len(raw))) + list(range(cdir, min(cdir + 200, eocd))) + list(range(pkl, pkl + 400)) +
             [int(v) for v in rs.randint(0, len(raw), 400)])
    survived = 0
    for pos in spots:
        for val in (0x00, 0xff, raw[pos] ^ 0x5a):
            b = bytearray(raw)
            b[pos] = val
            bad.write_bytes(bytes(b))
            n = lib.spb200_checkpoint_num_tensors(str(bad).encode())
            assert n < 0 or n <= 3
            if n >= 0:                                          # still parses: every tensor must still be readable or refused
                shape = (ctypes.c_int64 * 8)()
                rank = ctypes.c_int()
                buf = np.empty(64, np.float32)
                lib.spb200_checkpoint_tensor(str(bad).encode(), b'a.weight', ctypes.c_void_p(buf.ctypes.data), buf.size, shape,
                                             ctypes.byref(rank))
                survived += 1
    assert survived > 0
    # hostile sizes in the central directory: 0xFFFFFFFF sizes with a bogus ZIP64 extra field
    b = bytearray(raw)
    b[cdir + 20:cdir + 28] = b'\xff' * 8
    bad.write_bytes(bytes(b))
    assert lib.spb200_checkpoint_num_tensors(str(bad).encode()) < 0


def test_torchscript_archive_is_refused_with_a_clear_message(lib, tmp_path):
    """InferenceWrapper.trace writes <name>_script.pt next to <name>_params.pt (python/src/inferencewrapper.py:83-91);
    the script is a TorchScript archive, which this engine does not execute: the loader must say so."""
    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(1, 2, 3)

        def forward(self, x):
            return self.conv(x)

    p = str(tmp_path / 'tiny_script.pt')
    torch.jit.trace(Tiny(), torch.zeros(1, 1, 8, 8)).save(p)
    assert lib.spb200_checkpoint_num_tensors(p.encode()) < 0
    msg = lib.spb200_last_error(None).decode()
    assert 'TorchScript' in msg and '_params.pt' in msg


def test_dropin_module_initialize_descriptor_and_custom_ops_registered():
    """SuperPoint.initialize_descriptor (python/src/superpoint.py:86-89) resets the descriptor head's resettable direct
    children only; torch.ops.spb200.{forward,detect,detect_u8} are registered with fake kernels (shape inference without
    a GPU)."""
    import spb200
    net = spb200.SuperPoint(spb200.SuperPointSettings())
    up, bn, blk = net.descriptor.up_sample.weight.clone(), net.descriptor.bn.weight.clone(), net.descriptor.layer_in[0].conv1.weight.clone()
    with torch.no_grad():
        net.descriptor.bn.weight.fill_(3.0)
    net.initialize_descriptor()
    assert not torch.equal(up, net.descriptor.up_sample.weight)
    assert torch.equal(net.descriptor.bn.weight, torch.ones_like(bn))
    assert torch.equal(blk, net.descriptor.layer_in[0].conv1.weight)
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        img = torch.empty((2, 1, 240, 320), device='cuda')
        prob, desc, logits = torch.ops.spb200.forward(img, 1)
        assert prob.shape == (2, 240, 320) and desc.shape == (2, 128, 30, 40) and logits.shape == (2, 65, 30, 40)
        count, xy, conf, dsc = torch.ops.spb200.detect(img, 1, 500)
        assert count.shape == (2,) and xy.shape == (2, 500, 2) and dsc.shape == (2, 500, 128) and count.dtype == torch.int32
