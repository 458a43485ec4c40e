"""The wide parity set (tests/golden/forward_wide.npz, outputs of the reference made by make_golden_wide.py)."""
import os
import zlib

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
_wide = None


def wide():
    global _wide
    if _wide is None:
        _wide = np.load(os.path.join(GOLDEN, 'forward_wide.npz'))
    return _wide


def wide_cases(tag, h=None, fam=None):
    """Names ('shapes240_3', ...) of the cases stored for checkpoint tag 'm' (moderate) or 'h' (harsh)."""
    names = sorted(k[2:-4] for k in wide().files if k.startswith(tag + '_') and k.endswith('_crc'))
    if h is not None:
        names = [n for n in names if n.rsplit('_', 1)[0].endswith(str(h))]
    if fam is not None:
        names = [n for n in names if n.startswith(fam)]
    return names


def wide_image(name):
    """The seeded input image of a case; asserts that it is byte-identical to the one the reference saw."""
    from spb200 import synth
    fam, idx = name.rsplit('_', 1)
    kind = 'shapes' if fam.startswith('shapes') else 'rand'
    h = int(fam[len(kind):])
    w = {240: 320, 480: 640, 1088: 1920}[h]
    gray = synth.shapes_image(100 + int(idx), h, w) if kind == 'shapes' else synth.rand_image(100 + int(idx), h, w)
    return gray


def wide_ref(tag, name):
    g = wide()
    key = '%s_%s' % (tag, name)
    return {f: g['%s_%s' % (key, f)] for f in ('crc', 'xy', 'conf', 'cell', 'hsum', 'lsum', 'dsum', 'desc')}


def image_matches(gray, ref):
    return np.uint32(zlib.crc32(gray.numpy().tobytes())) == ref['crc']


def cell_max(prob, h, w):
    return np.asarray(prob, dtype=np.float32).reshape(h // 8, 8, w // 8, 8).max(axis=(1, 3))
