#!/usr/bin/env python
"""Benchmark of the SuperPoint inference hot path on B200 (contract: see the repository task description).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (network -> heatmap -> NMS -> sort -> descriptors, i.e.
InferenceWrapper.run of the reference, python/src/inferencewrapper.py:29-46, for every image) over one
batch of synthetic 480x640 grayscale images per GPU (BASELINE.json configs[2]; weak scaling for N > 1:
the batch per GPU is fixed, images are sharded across ranks with no collective).
Prints ONE JSON line on rank 0.

Other BASELINE configurations:
    configs[1]  python bench.py --detector-only --batch 32 --height 240 --width 320
    configs[3]  torchrun ... bench.py --gpus N --total-batch 512            (strong scaling: 512 / N images per GPU)
    configs[4]  torchrun ... bench.py --gpus 8 --height 1088 --width 1920 --top-k 2048 --batch 16
Split-precision levels (parity with margin): --precision fp16+layer1 | fp16+encoder | fp16+all.
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
    """SM clock, power and throttle reasons polled DURING the timed region: in-process NVML every ~2 ms (a 100 ms
    nvidia-smi loop cannot see a 30 ms region), nvidia-smi as the fallback."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.mode, self.sm_max = index, [], False, None, None, None

    def _nvml_loop(self, nv, h):
        while not self.stop_flag:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons')
                                  else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            try:
                h = nv.nvmlDeviceGetHandleByUUID(('GPU-' + uuid).encode())
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.mode = 'nvml'
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        try:
            q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = 'smi'
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
            self.thread.start()
        except Exception:
            self.mode = None

    def _smi_loop(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(',')]
            try:
                mask = sum(bit for bit, v in zip((0x8, 0x40, 0x20, 0x4), r[3:7]) if v.lower().startswith('active'))
                self.rows.append((float(r[0]), float(r[2]), mask))
                self.sm_max = float(r[1])
            except Exception:
                pass

    def stop(self):
        self.stop_flag = True
        if self.mode == 'smi':
            self.proc.terminate()
        if self.thread:
            self.thread.join(timeout=2)
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': self.sm_max, 'reasons': ['no samples' if self.mode else 'NVML and nvidia-smi unavailable']}
        sm = [r[0] for r in self.rows]
        pw = [r[1] for r in self.rows]
        mask = 0
        for r in self.rows:
            mask |= int(r[2])
        # "under load": samples with power in the upper half of what was seen
        thr = (max(pw) + min(pw)) / 2
        load = [c for c, p in zip(sm, pw) if p >= thr] or sm
        return {'sm_mhz': float(np.median(load)), 'sm_max_mhz': float(self.sm_max) if self.sm_max else None,
                'reasons': sorted(n for bit, n in self.REASONS.items() if mask & bit), 'power_w_max': max(pw), 'samples': len(sm),
                'source': self.mode}


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and so its pinned-memory allocations, first touch) to the CPUs of the GPU's NUMA node: with one
    process per GPU the host buffers of a rank then sit next to its PCIe root port.  Returns a note for the JSON line."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = '/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node' % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return 'numa node unknown'
        cpus = []
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return 'bound to NUMA node %d (%d cpus)' % (node, len(allowed))
        return 'numa node %d has no allowed cpus' % node
    except Exception as ex:
        return 'not bound (%s)' % type(ex).__name__


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
def time_e2e(eng, host_np, host_outs, cap, steps, pipelined):
    """Whole batches through the host-buffer C ABI: pinned frames in, pinned count / xy / conf / descriptors out, every
    step.  pipelined: spb200_detect_host_submit / _wait with two batches in flight (the call pattern of a caller that
    streams frames); otherwise one blocking spb200_detect_host per step.  Returns (seconds, keypoints downloaded)."""
    n_rot = len(host_np)
    kp = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if pipelined:
        ticket = eng.detect_host_submit(host_np[0], cap, out=host_outs[0])
        for i in range(steps):
            nxt = eng.detect_host_submit(host_np[(i + 1) % n_rot], cap, out=host_outs[(i + 1) % 2]) if i + 1 < steps else None
            out = eng.detect_host_wait(ticket)
            kp += int(out[0].sum())
            ticket = nxt
    else:
        for i in range(steps):
            out = eng.detect_host_u8(host_np[i % n_rot], cap, out=host_outs[0]) if host_np[0].dtype == np.uint8 else \
                eng.detect_host(host_np[i % n_rot], cap, out=host_outs[0])
            kp += int(out[0].sum())
    torch.cuda.synchronize()
    return time.perf_counter() - t0, kp


def measure_link(h2d_bytes, d2h_bytes, reps=6, sync=None, median=False):
    """What the host link of this box moves: a step's input bytes up and a step's output bytes down from / to pinned memory,
    each direction alone and both at once (two streams, CUDA events), as one copy and as four quarter copies (the
    pipeline's chunks), best of `reps` and of the two shapes.  The host-buffer path can be no faster than the download of
    its results while the next upload runs - `e2e.link` states how close it is.  With several ranks every repetition starts
    behind a barrier (`sync`) and the MEDIAN is reported: the ranks share the host's PCIe / memory fabric, and the best
    repetition of a rank is the one in which its neighbours happened to be idle."""
    h_in = torch.zeros(int(h2d_bytes), dtype=torch.uint8).pin_memory()
    h_out = torch.zeros(int(d2h_bytes), dtype=torch.uint8).pin_memory()
    d_in = torch.zeros(int(h2d_bytes), dtype=torch.uint8, device='cuda')
    d_out = torch.zeros(int(d2h_bytes), dtype=torch.uint8, device='cuda')
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def pieces(n, k):
        q = (n + k - 1) // k
        return [(i * q, min(n, (i + 1) * q)) for i in range(k) if i * q < n]

    def run(up, dn):
        ups, dns = [], []
        for r in range(reps):
            k = 1 if r % 2 == 0 else 4
            torch.cuda.synchronize()
            if sync is not None:
                sync()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if up:
                with torch.cuda.stream(s_up):
                    e[0].record()
                    for lo, hi in pieces(int(h2d_bytes), k):
                        d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
                    e[1].record()
            if dn:
                with torch.cuda.stream(s_dn):
                    e[2].record()
                    for lo, hi in pieces(int(d2h_bytes), k):
                        h_out[lo:hi].copy_(d_out[lo:hi], non_blocking=True)
                    e[3].record()
            torch.cuda.synchronize()
            if up:
                ups.append(e[0].elapsed_time(e[1]))
            if dn:
                dns.append(e[2].elapsed_time(e[3]))
        pick = (lambda v: float(np.median(v))) if median else min
        return (h2d_bytes / (pick(ups) * 1e-3) / 1e9 if up else None, d2h_bytes / (pick(dns) * 1e-3) / 1e9 if dn else None)

    run(True, True)
    up_alone, _ = run(True, False)
    _, dn_alone = run(False, True)
    up_both, dn_both = run(True, True)
    return {'h2d_gbs_alone': up_alone, 'd2h_gbs_alone': dn_alone, 'h2d_gbs_duplex': up_both, 'd2h_gbs_duplex': dn_both,
            'how': 'a step\'s bytes per direction between pinned host memory and the device, alone and both directions at once, '
                   'as one copy and as four, %s of %d (CUDA events); the duplex upload is half as long as the download, as in the pipeline'
                   % ('median' if median else 'best', reps)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='images per GPU per step')
    ap.add_argument('--total-batch', type=int, default=0, help='strong scaling: images per step over ALL GPUs (BASELINE configs[3]: 512)')
    ap.add_argument('--height', type=int, default=480)
    ap.add_argument('--width', type=int, default=640)
    ap.add_argument('--precision', default='fp16', help="fp32 | fp16 | bf16, optionally with split-precision stages: fp16+layer1 | fp16+encoder | fp16+all")
    ap.add_argument('--top-k', type=int, default=0)
    ap.add_argument('--detector-only', action='store_true', help='MagicPoint: heatmap + NMS, descriptor head skipped (BASELINE configs[1])')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the e2e variants, the next rows and the split-level table')
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
    numa_note = bind_to_gpu_numa_node(local_rank)            # before any pinned allocation
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    import spb200
    from spb200 import shard
    peaks = load_peaks()
    dev = torch.device('cuda', local_rank)
    H, W = args.height, args.width
    strong = args.total_batch > 0
    if strong:
        lo, hi = shard.shard_range(args.total_batch, rank, world)
        B = hi - lo
    else:
        B = args.batch

    eng = spb200.Engine(local_rank)
    eng.load_checkpoint(CKPT)
    eng.finalize(args.precision)
    eng.set_params(top_k=args.top_k, descriptor_enabled=not args.detector_only)
    cap = eng.max_keypoints(H, W) if not args.top_k else args.top_k

    n_rot = max(2, min(4, int(np.ceil(160e6 / (B * H * W * 4)))))        # rotating inputs larger than the 126 MB L2
    host_batches = make_batches(n_rot, B, H, W, rank)
    dev_batches = [b.to(dev) for b in host_batches]
    outs = eng.alloc_outputs(B, cap, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ---------------------------------------------------------
    # on a stream of its own: a call repeated with the same buffers is then replayed as one CUDA graph (the engine cannot
    # capture on the legacy default stream); the events are recorded on that stream
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        for i in range(max(args.warmup, 2 * n_rot)):        # two sights of every rotating batch: eager, then captured
            eng.detect(dev_batches[i % n_rot], cap, out=outs)
    barrier()
    eng.reset_kernel_launches()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record()
        for i in range(args.steps):
            eng.detect(dev_batches[i % n_rot], cap, out=outs)
        ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launches
    clocks = sampler.stop() if rank == 0 else None
    ms_max = shard.max_over_ranks(ms, dev)                  # device time, MAX over ranks
    value = shard.whole_job_rate(B * args.steps, ms * 1e-3, dev)
    kp_mean = float(outs[0].float().mean().item())

    # ---- end to end through the host-buffer C ABI ------------------------------------------------------
    # inputs and outputs live in PINNED host memory; every step uploads its frames and downloads count / xy / conf /
    # descriptors of the keypoints found.  Headline: the streaming form (submit / wait, two batches in flight, fp32
    # descriptors - the reference's output type); the blocking one-call form and the lighter formats are reported beside it.
    host_np = [b.pin_memory().numpy() for b in host_batches]
    host_outs = [eng.host_outputs(B, cap, True, pinned=True) for _ in range(2)]
    e2e_steps = max(3, min(args.steps, 20))
    time_e2e(eng, host_np, host_outs, cap, 3, True)
    barrier()
    dt, kp = time_e2e(eng, host_np, host_outs, cap, e2e_steps, True)
    e2e_value = shard.whole_job_rate(B * e2e_steps, dt, dev)
    d2h = kp * (8 + 4 + 512) // e2e_steps + 4 * B
    e2e_more = None
    link = None
    if not args.no_extras:
        barrier()
        link = measure_link(B * H * W * 4, max(int(d2h), 1), sync=barrier if world > 1 else None, median=world > 1)
        # the download of a step's results is the longest stage of the pipeline: its share of the duplex download rate
        link['d2h_gbs_e2e'] = (d2h * e2e_steps / dt) / 1e9         # this rank's own download rate inside the e2e loop
        link['e2e_frac_of_duplex_d2h'] = link['d2h_gbs_e2e'] / link['d2h_gbs_duplex']
        link['e2e_frac_of_d2h_alone'] = link['d2h_gbs_e2e'] / link['d2h_gbs_alone']
    if not args.no_extras:
        e2e_more = {}
        barrier()
        dt1, _ = time_e2e(eng, host_np, host_outs, cap, e2e_steps, False)
        e2e_more['blocking_call'] = {'value': shard.whole_job_rate(B * e2e_steps, dt1, dev), 'unit': UNIT, 'api': 'spb200_detect_host'}
        if base_prec != 'fp32' and not args.detector_only:
            u8 = [(b.squeeze(1) * 255).round().to(torch.uint8).pin_memory().numpy() for b in host_batches]
            barrier()
            time_e2e(eng, u8, host_outs, cap, 2, True)
            dt2, _ = time_e2e(eng, u8, host_outs, cap, e2e_steps, True)
            e2e_more['u8_frames'] = {'value': shard.whole_job_rate(B * e2e_steps, dt2, dev), 'unit': UNIT, 'h2d_bytes_per_step': B * H * W,
                                     'api': 'spb200_detect_host_submit(img_is_u8=1) / _wait'}
            eng.set_descriptor_format('fp16')
            outs16 = [eng.host_outputs(B, cap, True, pinned=True) for _ in range(2)]
            barrier()
            time_e2e(eng, u8, outs16, cap, 2, True)
            dt3, kp3 = time_e2e(eng, u8, outs16, cap, e2e_steps, True)
            e2e_more['u8_frames_fp16_descriptors'] = {'value': shard.whole_job_rate(B * e2e_steps, dt3, dev), 'unit': UNIT,
                                                      'h2d_bytes_per_step': B * H * W, 'd2h_bytes_per_step': kp3 * (8 + 4 + 256) // e2e_steps + 4 * B,
                                                      'api': 'spb200_set_descriptor_format(SPB200_DESC_FP16) + submit / wait'}
            time_e2e(eng, host_np, outs16, cap, 2, True)
            dt4, kp4 = time_e2e(eng, host_np, outs16, cap, e2e_steps, True)
            e2e_more['fp16_descriptors'] = {'value': shard.whole_job_rate(B * e2e_steps, dt4, dev), 'unit': UNIT,
                                            'd2h_bytes_per_step': kp4 * (8 + 4 + 256) // e2e_steps + 4 * B}
            eng.set_descriptor_format('fp32')
            if not args.top_k:
                eng.set_params(top_k=2048, descriptor_enabled=True)
                outs_k = [eng.host_outputs(B, 2048, True, pinned=True) for _ in range(2)]
                barrier()
                time_e2e(eng, host_np, outs_k, 2048, 2, True)
                dt5, kp5 = time_e2e(eng, host_np, outs_k, 2048, e2e_steps, True)
                e2e_more['top_k_2048'] = {'value': shard.whole_job_rate(B * e2e_steps, dt5, dev), 'unit': UNIT,
                                          'd2h_bytes_per_step': kp5 * (8 + 4 + 512) // e2e_steps + 4 * B}
                eng.set_params(top_k=args.top_k, descriptor_enabled=not args.detector_only)

    # ---- per-kernel CUDA-event profile (separate pass; roofline) -----------------------------------
    # every launch is bracketed by its own event pair on the launching stream: the kernels run one at a time with idle
    # gaps in between, i.e. at burst clocks - so the denominator is the BURST GEMM rate; the sustained one is printed too
    prof_steps = 8
    for i in range(2):
        eng.detect(dev_batches[i % n_rot], cap, out=outs)
    eng.profile_begin()
    for i in range(prof_steps):
        eng.detect(dev_batches[i % n_rot], cap, out=outs)
    entries = eng.profile_end()
    n_kp = float(outs[0].sum().item())
    esz = 4 if base_prec == 'fp32' else 2
    Hc, Wc = H // 8, W // 8
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
            # minimal bytes: the descriptor map read once + keypoint coordinates in + unit vectors out (corner re-reads hit L2)
            by = B * Hc * Wc * 128.0 * esz + n_kp * (8 + 512)
        row = {'kernel': name, 'ms': kms, 'share': kms / total_ms if total_ms else 0.0}
        if a['flops']:
            row['tflops'] = a['flops'] / (kms * 1e-3) / 1e12
            row['frac_tc_burst'] = row['tflops'] / peaks['tc_burst']
            row['frac_tc_sustained'] = row['tflops'] / peaks['tc_sustained']
        if by:
            row['gbs'] = by / (kms * 1e-3) / 1e9
            row['frac_hbm'] = row['gbs'] / peaks['hbm_gbs']
        table.append(row)
    table.sort(key=lambda r: -r['ms'])
    top = table[0]
    ridge = peaks['tc_burst'] * 1e12 / (peaks['hbm_gbs'] * 1e9)
    top_ai = (agg[top['kernel']]['flops'] / agg[top['kernel']]['bytes']) if agg[top['kernel']]['flops'] and agg[top['kernel']]['bytes'] else None
    if 'tflops' in top and not (top_ai is not None and top_ai < ridge):
        roofline = {'bound': 'tensor', 'kernel': top['kernel'], 'achieved': top['tflops'], 'peak': peaks['tc_burst'],
                    'unit': 'TFLOP/s', 'frac': top['tflops'] / peaks['tc_burst'], 'frac_of_sustained_peak': top['tflops'] / peaks['tc_sustained'],
                    'traffic': None}
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
    roofline['peak_source'] = peaks['source'] + (': burst bf16 GEMM rate (the kernels of the profile pass run one at a time); frac_of_sustained_peak uses the sustained rate'
                                                 if roofline['bound'] == 'tensor' else '')
    roofline['share_of_step'] = top['share']

    def conv_group(pred):
        rows = [(n_, a) for n_, a in agg.items() if pred(n_)]
        gms = sum(a['ms'] / a['n'] for _, a in rows)
        gfl = sum(a['flops'] for _, a in rows if a['flops'])
        if not gms or not gfl:
            return None
        tf = gfl / (gms * 1e-3) / 1e12
        return {'tflops': tf, 'frac_burst': tf / peaks['tc_burst'], 'frac_sustained': tf / peaks['tc_sustained'], 'ms': gms,
                'algorithmic_gflop_per_image': gfl / B / 1e9}
    roofline['all_convs'] = conv_group(lambda n_: agg[n_]['flops'] or n_ == 'image_planes')
    # the north star's "encoder convs": stem (with its image-plane pass) + layer1 + layer2
    roofline['encoder_convs'] = conv_group(lambda n_: n_ in ('image_planes', 'stem_pool') or n_.startswith('encoder.'))
    roofline['step_fraction_of_sustained_peak'] = (sum(a['flops'] for a in agg.values() if a['flops']) / (ms_max / args.steps * 1e-3) / 1e12 /
                                                   peaks['tc_sustained'])
    if args.profile_out and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        json.dump({'per_kernel': table, 'step_ms_profiled': total_ms, 'keypoints_per_image': kp_mean}, open(args.profile_out, 'w'), indent=1)

    # ---- the rows next to the path, rank 0 at N = 1 only, a few iterations each -------------------------
    next_rows = None
    if rank == 0 and world == 1 and base_prec != 'fp32' and not args.no_extras:
        next_rows = {}
        # N2: mutual-nearest-neighbour matching of consecutive images of the batch (B / 2 pairs per call)
        half = B // 2
        if half and not args.detector_only:
            eng.detect(dev_batches[0], cap, out=outs)
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
        if B >= hb and H >= hh and W >= hw:
            himg = dev_batches[0][:hb, :, :hh, :hw].contiguous()
            hs = hg.sample_homographies((hh, hw), cfg, np.random.default_rng(0))
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
            # N3: the C++ demo's loader on the device (resize + gray) in front of spb200_detect_u8, 720p BGR frames
            fr = torch.randint(0, 256, (hb, 720, 1280, 3), dtype=torch.uint8, device=dev)
            eng2.preprocess_u8(fr, hh, hw)
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for i in range(10):
                eng2.preprocess_u8(fr, hh, hw)
            p1.record()
            torch.cuda.synchronize()
            pms = p0.elapsed_time(p1) / 10
            next_rows['frame_loader_u8'] = {'value': hb / (pms * 1e-3), 'unit': 'frames/s', 'ms_per_call': pms, 'frames': '32 x 720x1280 BGR -> 240x320 gray',
                                            'api': 'spb200_preprocess_u8'}
            eng2.close()
        # the split-precision levels (parity with margin): device-resident throughput, few steps
        if not args.precision.count('+') and not args.detector_only:
            modes = {}
            for mode in ('fp16+layer1', 'fp16+encoder', 'fp16+all'):
                e3 = spb200.Engine(local_rank)
                e3.load_checkpoint(CKPT)
                e3.finalize(mode)
                e3.set_params(top_k=args.top_k)
                for i in range(3):
                    e3.detect(dev_batches[i % n_rot], cap, out=outs)
                torch.cuda.synchronize()
                q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                q0.record()
                for i in range(20):
                    e3.detect(dev_batches[i % n_rot], cap, out=outs)
                q1.record()
                torch.cuda.synchronize()
                modes[mode] = {'value': B * 20 / (q0.elapsed_time(q1) * 1e-3), 'unit': UNIT}
                e3.close()
            next_rows['split_precision_levels'] = modes

    # ---- CPU baseline (rank 0, N = 1 only, bounded sample) -----------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rate, n, secs = cpu_reference_rate(list(host_batches[0][:16]), 12.0, cores)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': '%d images of %dx%d in %.1f s, one at a time (the reference path is batch-1), torch fp32 CPU + C NMS' % (n, H, W, secs),
               'note': 'the oracle port: the reference itself imports in the build container but /root/reference does not exist on the GPU box'}

    if rank == 0:
        if args.detector_only:
            workload = 'magic_point-style detector only (heatmap + NMS, descriptor head skipped), batch %d per GPU at %dx%d grayscale (BASELINE configs[1])' % (B, H, W)
        elif strong:
            workload = 'super_point.pt keypoints+descriptors, %d images per step sharded over %d GPUs (%d per GPU) at %dx%d grayscale (BASELINE configs[3])' % (args.total_batch, world, B, H, W)
        elif H == 1088 and W == 1920:
            workload = 'super_point.pt keypoints+descriptors, batch %d per GPU at 1088x1920, top-k %d (BASELINE configs[4])' % (B, args.top_k)
        else:
            workload = 'super_point.pt keypoints+descriptors, batch %d per GPU at %dx%d grayscale (BASELINE configs[2])' % (B, H, W)
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'strong' if strong else 'weak', 'vs_baseline': None,
            'dtype': {'fp32': 'f32', 'fp16': 'f16', 'bf16': 'bf16'}[base_prec], 'data': 'synthetic',
            'config': {'workload': workload,
                       'batch_per_gpu': B, 'height': H, 'width': W, 'top_k': args.top_k, 'precision': args.precision,
                       'parallelism': 'batch-sharded x%d, no collective' % world,
                       'keypoints_per_image': kp_mean, 'weights': 'tests/golden/super_point.pt (synthetic recipe, reference-written)',
                       'l2': 'inputs rotate over %d batches (%.0f MB > 126 MB L2); activations are rewritten every step' % (n_rot, n_rot * B * H * W * 4 / 1e6),
                       'host': numa_note},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': B * H * W * 4, 'd2h_bytes_per_step': int(d2h),
                    'api': 'spb200_detect_host_submit / spb200_detect_host_wait: pinned host buffers in and out, two batches in flight '
                           '(upload / compute / download pipelined in 16-image chunks), fp32 descriptors; every step ends with its results on the host',
                    'steps': e2e_steps, 'link': link, 'variants': e2e_more},
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
