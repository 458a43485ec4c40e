"""First check of the CTA-pair kernel (SPB200_PAIR=1, halo_pair_kernel): the same batch through the default path and
through the paired path, outputs compared.  The two kernels do the same arithmetic in the same order, so the network
outputs should be identical.  Run on the GPU box under a timeout (the paired kernel has unbounded barrier waits):
    timeout -s KILL 60 python scripts/pair_check.py [batch height width]
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'feature-point-cnn_b200'))
import torch
import spb200
CKPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'super_point.pt')
b, h, w = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (2, 240, 320)
img = torch.rand((b, 1, h, w), generator=torch.Generator().manual_seed(1)).cuda()
outs = []
for pair in ('', '1'):
    if pair:
        os.environ['SPB200_PAIR'] = '1'
    else:
        os.environ.pop('SPB200_PAIR', None)
    e = spb200.Engine(0); e.load_checkpoint(CKPT); e.finalize('fp16'); e.set_params()
    prob, desc, logits = e.forward(img)
    torch.cuda.synchronize()
    outs.append((prob.cpu(), desc.cpu(), logits.cpu()))
    print('pair' if pair else 'default', 'ran: heatmap max %.4f' % float(prob.max()), flush=True)
    e.close()
for name, a, c in zip(('heatmap', 'descriptor map', 'logits'), outs[0], outs[1]):
    d = (a - c).abs()
    print('%-14s differing values %d of %d, max abs diff %.3e' % (name, int((a != c).sum()), a.numel(), float(d.max())))
