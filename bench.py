#!/usr/bin/env python
"""Benchmark of the SuperPoint inference hot path on B200 (contract: see the repository task description).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (network -> heatmap -> NMS -> sort -> descriptors, i.e.
InferenceWrapper.run of the reference, python/src/inferencewrapper.py:29-46, for every image) over one
batch of synthetic 480x640 grayscale images per GPU (BASELINE.json configs[2]; weak scaling for N > 1:
the batch per GPU is fixed, images are sharded across ranks with no collective).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))
CKPT = os.path.join(REPO, 'tests', 'golden', 'super_point.pt')
METRIC = 'SuperPoint images/sec @480x640 (kpts+desc)'
UNIT = 'images/s'


def ncu_traffic(kernel_substr, longest=False):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel, from the newest committed ncu
    full-set summary under profiles/ (MB columns of scripts/ncu_summary.py); None when there is none.  Several
    launches share a kernel name (one template instance serves several layers): `longest` picks the longest of them
    instead of the first."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(REPO, 'profiles', '*ncu_full_summary*.txt'))):
        cand = None
        for line in open(f):
            if kernel_substr in line:
                parts = line[44:].split()
                try:
                    us, val = float(parts[0]), (float(parts[1]) + float(parts[2])) * 1e6
                except Exception:
                    continue
                if cand is None or (longest and us > cand[0]):
                    cand = (us, val)
                if not longest:
                    break
        if cand is not None:
            best = cand[1]
    return best


def load_peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tc_burst': d['bf16_tflops'], 'tc_sustained': d['bf16_tflops_sustained'],
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'tc_burst': 1590.0, 'tc_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        # "under load": samples with power in the upper half of what was seen
        if sm:
            thr = (max(pw) + min(pw)) / 2 if pw else 0
            load = [s for s, p in zip(sm, pw) if p >= thr] or sm
            return {'sm_mhz': float(np.median(load)), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons),
                    'power_w_max': max(pw), 'samples': len(sm)}
        return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}


def make_batches(n_batches, batch, h, w, rank):
    """Synthetic 'shapes' images (spb200/synth.py), 16 distinct per batch, cycled; seeded per rank."""
    from spb200 import synth
    out = []
    for k in range(n_batches):
        base = torch.stack([synth.shapes_image(1000 * rank + 16 * k + i, h, w) for i in range(16)])
        reps = (batch + 15) // 16
        out.append(base.repeat(reps, 1, 1)[:batch, None].contiguous())
    return out


# --------------------------------------------------------------------------------------------------
# CPU arms (the oracle port of the reference's CPU path: this is the one place bench.py runs oracle/)
# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(images, max_seconds, threads):
    """InferenceWrapper.run per image (python/src/inferencewrapper.py:38-46) on the host cores -> images/s."""
    sys.path.insert(0, REPO)
    from oracle import postproc, weights
    torch.set_num_threads(threads)
    sd = weights.load_state_dict(CKPT)
    postproc.run(images[0][None], sd)          # warm-up (also builds the C NMS)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < max_seconds:
        postproc.run(images[n % len(images)][None], sd)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    imgs = make_batches(1, 16, args.height, args.width, 0)[0]
    per_step = 2
    sys.path.insert(0, REPO)
    from oracle import postproc, weights
    torch.set_num_threads(cores)
    sd = weights.load_state_dict(CKPT)
    for _ in range(max(args.warmup, 1)):
        postproc.run(imgs[0][None], sd)
    t0 = time.perf_counter()
    for s in range(args.steps):
        for j in range(per_step):
            postproc.run(imgs[(s * per_step + j) % 16][None], sd)
    dt = time.perf_counter() - t0
    value = args.steps * per_step / dt
    sample = '%d images of %dx%d per step, one at a time (the reference is batch-1), fp32 torch CPU + C NMS' % (
        per_step, args.height, args.width)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'super_point.pt keypoints+descriptors, 480x640 grayscale (BASELINE configs[2]), CPU arm: bounded sample',
                   'height': args.height, 'width': args.width},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='images per GPU per step')
    ap.add_argument('--height', type=int, default=480)
    ap.add_argument('--width', type=int, default=640)
    ap.add_argument('--precision', default='fp16', help="fp32 | fp16 | bf16, optionally with split-precision stages: fp16+layer1 | fp16+encoder | fp16+all")
    ap.add_argument('--top-k', type=int, default=0)
    ap.add_argument('--detector-only', action='store_true', help='MagicPoint: heatmap + NMS, descriptor head skipped (BASELINE configs[1])')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-out', default=None, help='write the per-kernel table (json) here')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    base_prec = args.precision.split('+')[0]

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))

    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    import spb200
    peaks = load_peaks()
    dev = torch.device('cuda', local_rank)
    B, H, W = args.batch, args.height, args.width

    eng = spb200.Engine(local_rank)
    eng.load_checkpoint(CKPT)
    eng.finalize(args.precision)
    eng.set_params(top_k=args.top_k, descriptor_enabled=not args.detector_only)
    cap = eng.max_keypoints(H, W) if not args.top_k else args.top_k

    n_rot = 4                                               # 4 x 78.6 MB of inputs > 126 MB L2
    host_batches = make_batches(n_rot, B, H, W, rank)
    dev_batches = [b.to(dev) for b in host_batches]
    outs = eng.alloc_outputs(B, cap, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ---------------------------------------------------------
    for i in range(args.warmup):
        eng.detect(dev_batches[i % n_rot], cap, out=outs)
    barrier()
    eng.reset_kernel_launches()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        eng.detect(dev_batches[i % n_rot], cap, out=outs)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launches
    clocks = sampler.stop() if rank == 0 else None
    from spb200 import shard
    ms_max = shard.max_over_ranks(ms, dev)                  # device time, MAX over ranks
    value = shard.whole_job_rate(B * args.steps, ms * 1e-3, dev)
    kp_mean = float(outs[0].float().mean().item())

    # ---- end to end through the host-buffer C ABI (spb200_detect_host) ------------------------------
    # inputs and outputs live in PINNED host memory; every step uploads its frames and downloads count / xy /
    # conf / descriptors of the keypoints found (the call pipelines upload, compute and download in chunks)
    host_np = [b.pin_memory().numpy() for b in host_batches]
    host_out = (torch.zeros((B,), dtype=torch.int32).pin_memory().numpy(),
                torch.zeros((B, cap, 2), dtype=torch.int32).pin_memory().numpy(),
                torch.zeros((B, cap), dtype=torch.float32).pin_memory().numpy(),
                torch.zeros((B, cap, 128), dtype=torch.float32).pin_memory().numpy())
    for i in range(3):
        host_out = eng.detect_host(host_np[i % n_rot], cap, out=host_out)
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    t0 = time.perf_counter()
    d2h = 0
    for i in range(e2e_steps):
        host_out = eng.detect_host(host_np[i % n_rot], cap, out=host_out)
        d2h += int(host_out[0].sum()) * (8 + 4 + 512) + 4 * B
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    e2e_value = shard.whole_job_rate(B * e2e_steps, dt, dev)

    # ---- per-kernel CUDA-event profile (separate pass; roofline) -----------------------------------
    prof_steps = 3
    eng.profile_begin()
    for i in range(prof_steps):
        eng.detect(dev_batches[i % n_rot], cap, out=outs)
    entries = eng.profile_end()
    n_kp = float(outs[0].sum().item())
    esz = 4 if base_prec == 'fp32' else 2
    agg = {}
    for name, kms, fl, by in entries:
        a = agg.setdefault(name, {'ms': 0.0, 'flops': fl, 'bytes': by, 'n': 0})
        a['ms'] += kms
        a['n'] += 1
    table = []
    total_ms = sum(a['ms'] for a in agg.values()) / prof_steps
    for name, a in agg.items():
        kms = a['ms'] / a['n']
        by = a['bytes']
        if name == 'nms_round0':
            by += 8.0 * n_kp             # survivor keys out
        if name == 'nms_finish_sort':
            by = 20.0 * n_kp             # survivor keys in, (x, y, conf) out
        if name == 'descriptors':
            by = n_kp * (4 * 128 * esz + 8 + 512)
        row = {'kernel': name, 'ms': kms, 'share': kms / total_ms if total_ms else 0.0}
        if a['flops']:
            row['tflops'] = a['flops'] / (kms * 1e-3) / 1e12
            row['frac_tc_sustained'] = row['tflops'] / peaks['tc_sustained']
        if by:
            row['gbs'] = by / (kms * 1e-3) / 1e9
            row['frac_hbm'] = row['gbs'] / peaks['hbm_gbs']
        table.append(row)
    table.sort(key=lambda r: -r['ms'])
    top = table[0]
    conv_rows = [r for r in table if 'tflops' in r]
    conv_ms = sum(r['ms'] for r in conv_rows)
    conv_flops = sum(a['flops'] for n_, a in agg.items() if a['flops'])
    # the bound of the dominant kernel: tensor pipe unless its algorithmic intensity is below the ridge of the
    # measured peaks (the fused stem: 131 FLOP/B vs a ridge of ~212), then HBM
    ridge = peaks['tc_sustained'] * 1e12 / (peaks['hbm_gbs'] * 1e9)
    top_ai = (agg[top['kernel']]['flops'] / agg[top['kernel']]['bytes']) if agg[top['kernel']]['flops'] and agg[top['kernel']]['bytes'] else None
    if 'tflops' in top and not (top_ai is not None and top_ai < ridge):
        roofline = {'bound': 'tensor', 'kernel': top['kernel'], 'achieved': top['tflops'], 'peak': peaks['tc_sustained'],
                    'unit': 'TFLOP/s', 'frac': top['tflops'] / peaks['tc_sustained'], 'traffic': None}
    else:
        roofline = {'bound': 'hbm', 'kernel': top['kernel'], 'achieved': top.get('gbs'), 'peak': peaks['hbm_gbs'],
                    'unit': 'GB/s', 'frac': top.get('frac_hbm'), 'traffic': None}
        if top_ai is not None:
            roofline['algorithmic_intensity_flop_per_byte'] = top_ai
            roofline['ridge_flop_per_byte'] = ridge
    ncu_name = {'stem_pool': 'stem_planes_kernel', 'nms_round0': 'nms_round0_kernel', 'nms_finish_sort': 'nms_finish_kernel', 'descriptors': 'sample_desc', 'heatmap': 'heatmap_kernel',
                'image_planes': 'planes_kernel'}.get(top['kernel'])
    roofline['traffic'] = ncu_traffic(ncu_name) if ncu_name and B == 64 and H == 480 and W == 640 else None
    if top['kernel'] == 'descriptor.layer_out.0' and B == 64 and H == 480 and W == 640:
        # the 256 -> 128 block with the concatenated input: the longest launch of this template instance in the capture
        roofline['traffic'] = ncu_traffic('halo_tc_kernel<128, 2, 1, 2, ', longest=True)
    roofline['traffic_source'] = 'ncu --set full capture of the same workload committed under profiles/ (dram bytes read + written per launch)'
    roofline['peak_source'] = peaks['source'] + (', sustained bf16 GEMM figure' if roofline['bound'] == 'tensor' else '')
    roofline['share_of_step'] = top['share']
    roofline['all_convs'] = {'tflops': conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms else None,
                             'frac': conv_flops / (conv_ms * 1e-3) / 1e12 / peaks['tc_sustained'] if conv_ms else None,
                             'ms': conv_ms, 'algorithmic_gflop_per_image': conv_flops / B / 1e9}
    if args.profile_out and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        json.dump({'per_kernel': table, 'step_ms_profiled': total_ms, 'keypoints_per_image': kp_mean}, open(args.profile_out, 'w'), indent=1)

    # ---- the rows next to the path, rank 0 at N = 1 only, a few iterations each -------------------------
    next_rows = None
    if rank == 0 and world == 1 and base_prec != 'fp32':
        next_rows = {}
        # N3: 8-bit frames through the host-buffer entry point (a quarter of the upload bytes)
        u8 = [(b.squeeze(1) * 255).round().to(torch.uint8).pin_memory().numpy() for b in host_batches]
        for i in range(2):
            host_out = eng.detect_host_u8(u8[i % n_rot], cap, out=host_out)
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            host_out = eng.detect_host_u8(u8[i % n_rot], cap, out=host_out)
        torch.cuda.synchronize()
        next_rows['e2e_u8_frames'] = {'value': B * e2e_steps / (time.perf_counter() - t0), 'unit': UNIT, 'h2d_bytes_per_step': B * H * W,
                                      'api': 'spb200_detect_host_u8'}
        # N2: mutual-nearest-neighbour matching of consecutive images of the batch (B / 2 pairs per call)
        half = B // 2
        if half:
            da, db, ca, cb = outs[3][:half], outs[3][half:2 * half], outs[0][:half], outs[0][half:2 * half]
            eng.match(da, ca, db, cb, 0.7)
            torch.cuda.synchronize()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            for i in range(5):
                mres = eng.match(da, ca, db, cb, 0.7)
            m1.record()
            torch.cuda.synchronize()
            mms = m0.elapsed_time(m1) / 5
            nq = float(ca.float().mean().item())
            next_rows['match'] = {'value': half / (mms * 1e-3), 'unit': 'image pairs/s', 'ms_per_call': mms, 'pairs_per_call': half,
                                  'keypoints_per_image': nq, 'matched_fraction': float((mres[0] >= 0).float().sum().item() / max(ca.sum().item(), 1)),
                                  'tflops_fp32': 2.0 * half * nq * nq * 128 / (mms * 1e-3) / 1e12, 'api': 'spb200_match'}
        # N1: homography adaptation, the reference's default configuration (15 homographies), batch 32 at 240x320
        from spb200 import homographies as hg
        cfg = hg.HomographyConfig()
        hb, hh, hw = 32, 240, 320
        himg = dev_batches[0][:hb, :, :hh, :hw].contiguous()
        hs = hg.sample_homographies((hh, hw), cfg, __import__('numpy').random.default_rng(0))
        eng2 = spb200.Engine(local_rank)                      # its own engine: the workspace is per (B, H, W)
        eng2.load_checkpoint(CKPT)
        eng2.finalize(args.precision)
        eng2.set_params()
        eng2.homography_adaptation(himg, hs, cfg.valid_border_margin, cfg.aggregation)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(5):
            eng2.homography_adaptation(himg, hs, cfg.valid_border_margin, cfg.aggregation)
        torch.cuda.synchronize()
        hdt = (time.perf_counter() - t0) / 5
        next_rows['homography_adaptation'] = {'value': hb / hdt, 'unit': 'images/s', 'ms_per_call': hdt * 1e3, 'batch': hb, 'height': hh,
                                              'width': hw, 'homographies': cfg.num, 'api': 'spb200_homography_adaptation'}
        eng2.close()

    # ---- CPU baseline (rank 0, N = 1 only, bounded sample) -----------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rate, n, secs = cpu_reference_rate(list(host_batches[0][:16]), 12.0, cores)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': '%d images of %dx%d in %.1f s, one at a time (the reference path is batch-1), torch fp32 CPU + C NMS' % (n, H, W, secs)}

    if rank == 0:
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'fp32': 'f32', 'fp16': 'f16', 'bf16': 'bf16'}[base_prec], 'data': 'synthetic',
            'config': {'workload': ('magic_point-style detector only (heatmap + NMS, descriptor head skipped), batch %d per GPU at %dx%d grayscale (BASELINE configs[1])' if args.detector_only else 'super_point.pt keypoints+descriptors, batch %d per GPU at %dx%d grayscale (BASELINE configs[2])') % (B, H, W),
                       'batch_per_gpu': B, 'height': H, 'width': W, 'top_k': args.top_k, 'precision': args.precision, 'parallelism': 'batch-sharded x%d, no collective' % world,
                       'keypoints_per_image': kp_mean, 'weights': 'tests/golden/super_point.pt (synthetic recipe, reference-written)',
                       'l2': 'inputs rotate over %d batches (%.0f MB > 126 MB L2); activations are rewritten every step' % (n_rot, n_rot * B * H * W * 4 / 1e6)},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': B * H * W * 4, 'd2h_bytes_per_step': d2h // e2e_steps,
                    'api': 'spb200_detect_host: pinned host buffers in/out, chunked upload / compute / download pipeline, returns when the results are on the host', 'steps': e2e_steps},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': roofline,
            'next_rows': next_rows,
            'cpu_baseline': cpu,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
