"""The detector tail with the fused softmax / depth-to-space epilogue against the logits path (SPB200_NO_FUSED_HEAT=1):
heatmap, counts, keypoints, confidences and descriptors must be bit-identical."""
import os, sys, subprocess
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch, spb200
    from spb200 import synth
    out = {}
    for (B, H, W) in [(3, 240, 320), (5, 480, 640), (1, 1088, 1920), (2, 112, 176)]:
        e = spb200.Engine(0); e.load_checkpoint(os.path.join(REPO, 'tests/golden/super_point.pt')); e.finalize('fp16'); e.set_params()
        img = torch.stack([synth.shapes_image(i, H, W) for i in range(B)])[:, None].contiguous().cuda()
        cap = e.max_keypoints(H, W)
        for rep in range(3):                                   # eager, first graph capture, graph replay
            count, xy, conf, desc, prob = e.detect(img, cap, want_prob=True)
        torch.cuda.synchronize()
        n = count.cpu()
        out[(B, H, W)] = dict(count=n, prob=prob.cpu(), xy=[xy[i, :int(n[i])].cpu() for i in range(B)],
                              conf=[conf[i, :int(n[i])].cpu() for i in range(B)], desc=[desc[i, :int(n[i])].cpu() for i in range(B)])
        for rep in range(3):                                   # without the heatmap: round 0 multiplies by the per-cell normaliser itself
            count2, xy2, conf2, desc2 = e.detect(img, cap)[:4]
        torch.cuda.synchronize()
        n2 = count2.cpu()
        assert torch.equal(n, n2), 'counts with and without the heatmap output differ'
        for i in range(B):
            k = int(n[i])
            assert torch.equal(xy[i, :k], xy2[i, :k]) and torch.equal(conf[i, :k], conf2[i, :k]) and torch.equal(desc[i, :k], desc2[i, :k]), \
                'keypoints with and without the heatmap output differ'
        e.close()
    torch.save(out, sys.argv[2])
    sys.exit(0)
import torch
res = []
for tag, env in (('fused', {}), ('plain', {'SPB200_NO_FUSED_HEAT': '1'})):
    path = '/tmp/fh_%s.pt' % tag
    subprocess.check_call([sys.executable, __file__, 'child', path], env={**os.environ, **env})
    res.append(torch.load(path))
ok = True
for key in res[0]:
    a, b = res[0][key], res[1][key]
    same = bool((a['count'] == b['count']).all()) and torch.equal(a['prob'], b['prob'])
    for k in ('xy', 'conf', 'desc'):
        same = same and all(torch.equal(x, y) for x, y in zip(a[k], b[k]))
    print(key, 'keypoints', a['count'].tolist(), 'identical' if same else 'DIFFERENT', 'max heat diff %.3g' % float((a['prob'] - b['prob']).abs().max()))
    ok = ok and same
print('fused heat check', 'OK' if ok else 'FAILED')
sys.exit(0 if ok else 1)
