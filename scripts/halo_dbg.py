"""Debug: timeline of one fused 128-channel halo kernel (CTA 0): MMA issuer and first epilogue warp."""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'feature-point-cnn_b200'))
import numpy as np, torch
os.environ['SPB200_HALO_DBG'] = sys.argv[1] if len(sys.argv) > 1 else '11'
import spb200
from spb200 import _lib
CKPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'super_point.pt')
b, h, w = 64, 480, 640
img = torch.rand((b, 1, h, w), generator=torch.Generator().manual_seed(1)).cuda()
e = spb200.Engine(0); e.load_checkpoint(CKPT); e.finalize('fp16'); e.set_params()
cap = e.max_keypoints(h, w)
for i in range(3): out = e.detect(img, cap)
torch.cuda.synchronize()
lib = e._lib
buf = (ctypes.c_longlong * 128)()
print('rc', lib.spb200_debug_halo(ctypes.cast(buf, ctypes.c_void_p)))
a = np.array(list(buf), dtype=np.int64).reshape(16, 8)
t0 = a[0, 0]
names = ['m:start', 'm:g1issued', 'm:yfull', 'm:g2issued', 'e:d1full', 'e:epi1done', 'e:d2full', 'e:epi2done']
for j in range(9):
    print(j, ' '.join('%s=%6d' % (n, a[j, k] - t0) for k, n in enumerate(names)))
