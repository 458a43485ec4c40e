"""The resident-weight blocks on CTA pairs (SPB200_PAIR64=1) against the single-CTA kernel: same batch through both (one
process each: the switch is read once), outputs compared; then per-kernel times.  Run under a timeout on the GPU box:
    timeout -s KILL 300 python scripts/pair64_check.py [batch height width]
"""
import os, subprocess, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
code = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, 'feature-point-cnn_b200'))
import torch, spb200
b, h, w = %d, %d, %d
img = torch.rand((b, 1, h, w), generator=torch.Generator().manual_seed(1)).cuda()
e = spb200.Engine(0); e.load_checkpoint(os.path.join(%r, 'tests', 'golden', 'super_point.pt')); e.finalize('fp16'); e.set_params()
prob, desc, logits = e.forward(img)
torch.cuda.synchronize()
l1a, l1b = e.export_activation('l1a', b), e.export_activation('l1b', b)
torch.save({'prob': prob.cpu(), 'desc': desc.cpu(), 'l1a': l1a.cpu(), 'l1b': l1b.cpu()}, sys.argv[1])
e.profile_begin()
for _ in range(3): e.forward(img)
t = {}
for name, ms, fl, by in e.profile_end(): t[name] = t.get(name, 0) + ms / 3
print('ran: heatmap max %%.4f  layer1.0 %%.4f ms  layer1.1 %%.4f ms' %% (float(prob.max()), t.get('encoder.layer1.0', 0), t.get('encoder.layer1.1', 0)), flush=True)
'''
b, h, w = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (2, 240, 320)
outs = []
for pair in ('0', '1'):
    env = dict(os.environ); env['SPB200_PAIR64'] = pair
    f = '/tmp/pair64_%s.pt' % pair
    r = subprocess.run([sys.executable, '-c', code % (REPO, b, h, w, REPO), f], env=env, capture_output=True, text=True, timeout=120)
    print('PAIR64=%s' % pair, r.stdout.strip() or r.stderr[-800:], flush=True)
    outs.append(f)
import torch
a, c = torch.load(outs[0]), torch.load(outs[1])
for k in a:
    d = (a[k] - c[k]).abs()
    print('%-6s differing values %d of %d, max abs diff %.3e' % (k, int((a[k] != c[k]).sum()), a[k].numel(), float(d.max())))
