"""Debug: run the plane-fed stem once and compare with the im2col stem."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'feature-point-cnn_b200'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import spb200
CKPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'super_point.pt')
b, h, w = 1, 240, 320
img = torch.rand((b, 1, h, w), generator=torch.Generator().manual_seed(1))
outs = []
for old in sys.argv[1:] or ['1', '0']:
    os.environ['SPB200_OLD_STEM'] = old
    e = spb200.Engine(0); e.load_checkpoint(CKPT); e.finalize('fp16'); e.set_params()
    e.forward(img.cuda()); torch.cuda.synchronize()
    outs.append(e.export_activation('pool', b).cpu()); e.close()
    print('old' if old == '1' else 'new', 'ok', float(outs[-1].abs().max()))
if len(outs) == 2:
    d = (outs[0] != outs[1])
    print('differ', int(d.sum()), 'of', d.numel(), 'max', float((outs[0] - outs[1]).abs().max()))
    if d.any():
        idx = torch.nonzero(d)
        print(idx[:10].tolist()); print('ys', sorted(set(idx[:, 2].tolist()))[:40]); print('xs', sorted(set(idx[:, 3].tolist()))[:40])
