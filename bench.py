"""Benchmark of the PNN hot path (BASELINE.json configs[1]).

Workload: 100 synthetic BSDS-shaped 320x480 luminance images, every W-aligned block with an in-image context
anchor, for W in {4, 8, 16, 32} (nets FC-4, FC-8, CONV-16, CONV-32, seeded random init -- the pretrained HM
weights are not shipped with the reference), outputs = rounded uint8 predictions, per-block PSNR (float64) and win
flag against a supplied baseline PSNR array; ONE gather of the statistics to rank 0.  A "step" is one pass over
all four block sizes.  `--scaling strong` (default, what BASELINE.json states: the 100 images are sharded round-robin
over the N GPUs) or `--scaling weak` (100 images per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling strong|weak]
    python bench.py --config conv64 | hm       (builder-run records of BASELINE.json configs[2] / configs[3])

`value`  : predictions/s with images and block lists resident in HBM (device-pointer C-ABI calls).
`e2e`    : the same metric through the host-pointer C-ABI call (pinned host buffers; H2D of the images
           and block lists and D2H of predictions + PSNRs inside the timed region).
`--impl reference`: the CPU stand-in of the reference path (fp32 oracle on torch-CPU, all host threads, the
           reference's batch size of 10) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WIDTHS = ((4, True), (8, True), (16, False), (32, False))
N_IMAGES, HEIGHT, WIDTH_IMAGE = int(os.environ.get('PNN_BENCH_IMAGES', '100')), 320, 480   # (the variable: test aid for ragged shards)
MEAN = 117.8952234192841
METRIC = 'PNN predictions/sec (4x4+8x8+16x16+32x32 blocks of 100 BSDS-shaped images)'
# algorithmic FLOPs per prediction = 2 * MACs (dense count)
MACS = {4: 2995200, 8: 3340800, 16: 48750592, 32: 273317888}


_JSON_FD = None


def emit(line):
    text = json.dumps(line) + '\n'
    if _JSON_FD is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, text.encode())


def synthetic_image(height, width, seed):
    """clip(128 + 60 sin(x/17) + 40 cos(y/11) + N(0, 4^2)), default_rng(seed)."""
    rng = numpy.random.default_rng(seed)
    y, x = numpy.mgrid[0:height, 0:width]
    img = 128. + 60. * numpy.sin(x / 17.) + 40. * numpy.cos(y / 11.) + rng.normal(0., 4., (height, width))
    return numpy.clip(numpy.round(img), 0, 255).astype(numpy.uint8)


def workload_config(n_gpus, scaling='strong'):
    """-> (config dict, blocks of the WHOLE job per width)."""
    from context_adaptive_neural_network_based_prediction_b200 import offline
    n_images = N_IMAGES if scaling == 'strong' else N_IMAGES * n_gpus
    blocks = {w: len(offline.grid_blocks(HEIGHT, WIDTH_IMAGE, w)[0]) * n_images for w, _ in WIDTHS}
    return {
        'workload': 'configs[1]: PNN 4x4/8x8/16x16/32x32 batched prediction over 100 synthetic BSDS-shaped '
                    '(320x480 after the reference 1-px crop) images, PSNR / win statistics gathered',
        'scaling_mode': 'strong: 100 images in total, image i on rank i % N (offline.shard_image_indices)' if scaling == 'strong'
                        else 'weak: 100 images per GPU',
        'images_total': n_images, 'image_shape': [HEIGHT, WIDTH_IMAGE],
        'nets': ['FC-4', 'FC-8', 'CONV-16', 'CONV-32'], 'weights': 'seeded random init (reference initialisers)',
        'blocks_total': {str(w): n for w, n in blocks.items()}, 'masks': [0, 0],
        'parallelism': 'images sharded over %d GPU(s), no data-path collective, one NCCL gather of statistics' % n_gpus,
        'l2': 'explicit flush (256 MiB write) before every timed step; per-step activations (GBs) exceed the 126 MB L2 anyway',
    }, blocks


class ClockSampler(object):
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index),
                 '--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
                 'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
                 'clocks_event_reasons.sw_power_cap', '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, sm_max, reasons = [], None, set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for t, line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7 or not (t0 - 0.1 <= t <= t1 + 0.3):
                continue
            try:
                sm.append(float(parts[0]))
                sm_max = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(numpy.median(sm)) if sm else None, 'sm_max_mhz': sm_max, 'samples': len(sm),
                'reasons': sorted(reasons)}


def make_weights(directory):
    from context_adaptive_neural_network_based_prediction_b200 import weights
    paths, tensors = {}, {}
    for w, is_fc in WIDTHS:
        wts = weights.init_weights(w, is_fc, seed=w)          # rng(W), reference initialisers
        path = os.path.join(directory, 'net_%d.pnnw' % w)
        weights.save_flat(path, w, is_fc, wts)
        paths[w], tensors[w] = path, wts
    return paths, tensors


# ----------------------------------------------------------------------------------------------
# CPU stand-in of the reference path (bounded sample)
# ----------------------------------------------------------------------------------------------

def cpu_reference_rates(tensors, seconds_per_size, batch_size=10):
    """fp32 oracle on torch-CPU with every host thread, driven like pnn/batching.py:74-87 (sess.run per batch of 10).

    Only the network forward is timed (contexts pre-gathered): generous to the CPU.
    """
    import torch
    from oracle import context, nets
    torch.set_num_threads(os.cpu_count() or 1)
    images = numpy.stack([synthetic_image(HEIGHT, WIDTH_IMAGE, s) for s in range(2)])
    from context_adaptive_neural_network_based_prediction_b200 import offline
    rates, counts = {}, {}
    for w, is_fc in WIDTHS:
        rows, cols = offline.grid_blocks(HEIGHT, WIDTH_IMAGE, w)
        n = min(len(rows), 400)
        above, left, flat, _ = context.gather_image_blocks(images, numpy.zeros(n, int), rows[:n], cols[:n], w, MEAN, 0, 0)
        nets.forward(tensors[w], w, is_fc, (flat[:batch_size],) if is_fc else (above[:batch_size], left[:batch_size]))
        done, t0 = 0, time.perf_counter()
        while True:
            i = (done // batch_size * batch_size) % (n - batch_size + 1)
            if is_fc:
                nets.forward_fc(tensors[w], flat[i:i + batch_size])
            else:
                nets.forward_conv(tensors[w], above[i:i + batch_size], left[i:i + batch_size])
            done += batch_size
            if time.perf_counter() - t0 >= seconds_per_size and done >= 3 * batch_size:
                break
        rates[w] = done / (time.perf_counter() - t0)
        counts[w] = done
    return rates, counts


def aggregate_rate(rates, blocks):
    total = float(sum(blocks.values()))
    return total / sum(blocks[w] / rates[w] for w in blocks)


def run_reference(args, rank):
    if rank != 0:
        return
    cfg, blocks = workload_config(args.gpus, args.scaling)
    tmp = tempfile.mkdtemp(prefix='pnn_bench_')
    _, tensors = make_weights(tmp)
    for _ in range(max(1, min(args.warmup, 3))):
        cpu_reference_rates(tensors, 0.3)                     # warm-up
    values, counts, step_seconds = [], {}, []
    t_all = time.perf_counter()
    for _ in range(max(1, args.steps)):
        t_step = time.perf_counter()
        rates, counts = cpu_reference_rates(tensors, args.reference_seconds)
        step_seconds.append(time.perf_counter() - t_step)
        values.append(aggregate_rate(rates, blocks))
    value = float(numpy.mean(values))
    cores = os.cpu_count() or 1
    sample = ('network forward only, batches of 10 (pnn/batching.py), %s blocks per size per step, rate extrapolated to the '
              'full block mix' % {str(k): v for k, v in counts.items()})
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'predictions/s', 'n_gpus': args.gpus,
        # a step of this arm is a BOUNDED SAMPLE of the workload (its measured wall time is `ms_per_step`); `value` is the
        # measured rate of the sample, weighted by the block mix of the full workload
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * float(numpy.mean(step_seconds)),
        'ms_full_workload_at_this_rate': 1e3 * sum(blocks.values()) / value,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': cfg,
        'cpu_baseline': {'value': value, 'unit': 'predictions/s', 'cores': cores, 'kind': 'port', 'sample': sample,
                         'note': 'TensorFlow 1.x is not installable offline; the fp32 oracle on torch-CPU (oneDNN/MKL) stands in '
                                 'for the reference TF-CPU path'},
        'e2e': {'value': value, 'unit': 'predictions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'wall_s': time.perf_counter() - t_all,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'],
                    help='strong (BASELINE.json configs[1]): 100 images in total sharded over the GPUs; weak: 100 images per GPU')
    ap.add_argument('--config', default='offline', choices=['offline', 'conv64', 'hm'],
                    help='offline = BASELINE.json configs[1] (the driver\'s bench); conv64 = configs[2]; hm = configs[3]')
    ap.add_argument('--reference-seconds', type=float, default=3.0, help='CPU seconds per block size per step')
    ap.add_argument('--cpu-baseline-seconds', type=float, default=4.0)
    ap.add_argument('--precision', default='bf16x3', choices=['bf16x3', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--skip-e2e', action='store_true', help='profiling runs only: skip the host-buffer arm')
    ap.add_argument('--report', action='store_true', help='print the per-kernel table to stderr')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    # libraries (NCCL's version banner) write to fd 1; keep stdout for the ONE JSON line
    sys.stdout.flush()
    global _JSON_FD
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)

    if args.config != 'offline':
        if rank == 0:
            import bench_configs
            (bench_configs.run_conv64 if args.config == 'conv64' else bench_configs.run_hm)(args, emit)
        return
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from context_adaptive_neural_network_based_prediction_b200 import Engine, offline

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    args.warmup = max(args.warmup, 3)
    cfg, blocks = workload_config(world, args.scaling)
    tmp = tempfile.mkdtemp(prefix='pnn_bench_')
    paths, tensors = make_weights(tmp)
    eng = Engine(mean_training=MEAN, device=local_rank)
    for w, _ in WIDTHS:
        eng.load_net(paths[w])
    eng.set_precision(args.precision)

    # this rank's image shard.  strong: image i of the 100 lives on rank i % N (13 or 12 images per rank at N = 8);
    # weak: images rank*100 .. rank*100+99 of a 100*N image set
    if args.scaling == 'strong':
        seeds_of = [offline.shard_image_indices(N_IMAGES, r, world) for r in range(world)]
    else:
        seeds_of = [list(range(r * N_IMAGES, (r + 1) * N_IMAGES)) for r in range(world)]
    n_local = len(seeds_of[rank])
    blocks_per_image = sum(len(offline.grid_blocks(HEIGHT, WIDTH_IMAGE, w)[0]) for w, _ in WIDTHS)
    counts = [len(sd) * blocks_per_image for sd in seeds_of]              # blocks per rank, known to everybody
    ragged = len(set(counts)) > 1
    images_np = numpy.stack([synthetic_image(HEIGHT, WIDTH_IMAGE, sd) for sd in seeds_of[rank]])
    images_pin = torch.from_numpy(images_np).pin_memory()
    d_images = images_pin.to(dev)
    total_job = sum(blocks.values())
    total = counts[rank]
    d_psnr = torch.empty(total, dtype=torch.float64, device=dev)
    d_win = torch.empty(total, dtype=torch.uint8, device=dev)
    d_base = torch.empty(total, dtype=torch.float64, device=dev)            # supplied baseline PSNRs: best HEVC intra mode
    h_psnr_all = torch.empty(total, dtype=torch.float64).pin_memory()      # e2e arm: the PSNRs of all sizes land in one pinned array
    h_gathered = torch.empty(sum(counts), dtype=torch.float64).pin_memory() if (rank == 0 and world > 1) else None
    per_w, off = {}, 0
    for w, is_fc in WIDTHS:
        idx, rows, cols = offline.blocks_of_images(n_local, HEIGHT, WIDTH_IMAGE, w)
        n = len(rows)
        per_w[w] = {
            'is_fc': is_fc, 'n': n, 'off': off,
            'h_idx': torch.from_numpy(idx).pin_memory(), 'h_rows': torch.from_numpy(rows).pin_memory(),
            'h_cols': torch.from_numpy(cols).pin_memory(),
            'd_u8': torch.empty((n, w * w), dtype=torch.uint8, device=dev),
            'h_u8': torch.empty((n, w, w), dtype=torch.uint8).pin_memory(),
            'h_psnr': h_psnr_all[off:off + n],
        }
        for k in ('idx', 'rows', 'cols'):
            per_w[w]['d_' + k] = per_w[w]['h_' + k].to(dev)
        off += n
    # the baseline PSNR array of the workload = PSNR of the best of the 35 HEVC intra modes per block (the reference's
    # psnrs_hevc_best_mode), computed once on the GPU OUTSIDE the timed region and timed on its own
    torch.cuda.synchronize()
    hb0, hb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hb0.record()
    for w, _ in WIDTHS:
        p = per_w[w]
        eng.hevc_best_mode_device(w, d_images.data_ptr(), n_local, HEIGHT, WIDTH_IMAGE, p['d_idx'].data_ptr(),
                                  p['d_rows'].data_ptr(), p['d_cols'].data_ptr(), p['n'], (0, 0), None,
                                  d_base.data_ptr() + 8 * p['off'], None, torch.cuda.current_stream().cuda_stream)
    hb1.record()
    torch.cuda.synchronize()
    hevc_baseline_ms = hb0.elapsed_time(hb1)
    base_host = d_base.cpu()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    size_events = {w: [] for w, _ in WIDTHS}

    def step_device(record=False):
        stream = torch.cuda.current_stream().cuda_stream
        for w, _ in WIDTHS:
            p = per_w[w]
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            eng.predict_image_blocks_device(w, p['is_fc'], d_images.data_ptr(), n_local, HEIGHT, WIDTH_IMAGE,
                                            p['d_idx'].data_ptr(), p['d_rows'].data_ptr(), p['d_cols'].data_ptr(), p['n'],
                                            (0, 0), None, p['d_u8'].data_ptr(), d_psnr.data_ptr() + 8 * p['off'], stream)
            if record:
                e1.record()
                size_events[w].append((e0, e1))
        eng.win_flags_device(d_psnr.data_ptr(), d_base.data_ptr(), total, d_win.data_ptr(), stream)
        return offline.gather_statistics(d_psnr, d_win, rank, world, counts=counts if ragged else None)

    def step_e2e():
        # smallest block list first: its host-side validation is short, so the GPU starts at once and the validation of
        # the larger lists overlaps the kernels
        for w, _ in sorted(WIDTHS, key=lambda wf: per_w[wf[0]]['n']):
            p = per_w[w]
            # asynchronous form of the host-pointer call: the four block sizes are enqueued back to back, uploads and
            # read-backs overlap the kernels of the neighbouring size; every output is read after synchronize()
            eng.predict_image_blocks(w, p['is_fc'], images_pin.numpy(), p['h_rows'].numpy(), p['h_cols'].numpy(),
                                     p['h_idx'].numpy(), want_float=False, out_uint8=p['h_u8'].numpy(),
                                     out_psnr=p['h_psnr'].numpy(), wait=False)
        eng.synchronize()
        psnr_all = h_psnr_all
        wins = psnr_all > base_host                      # = (psnr - baseline > 0), comparing_pnn_ipfcns_hevc_best_mode.py:87
        if world > 1:
            g_psnr, g_win = offline.gather_statistics(psnr_all.to(dev, non_blocking=True), wins.to(dev, non_blocking=True),
                                                      rank, world, counts=counts if ragged else None)
            if rank == 0:
                return offline.reduce_statistics_device(g_psnr, g_win, pinned=h_gathered)
            return None
        return offline.reduce_statistics(psnr_all.numpy(), wins.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, record_sizes=False):
        """K steps, each preceded by an L2 flush, timed with CUDA events on the launching stream."""
        events = []
        barrier()
        t0 = time.time()
        for _ in range(steps):
            flush_buf.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(True) if record_sizes else fn()
            e1.record()
            events.append((e0, e1))
        barrier()
        t1 = time.time()
        ms = sum(a.elapsed_time(b) for a, b in events)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t0, t1

    # ---- device-resident arm -------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = eng.launch_count
    eng.set_profiling(True)
    eng.profile_report()
    ms_total, t0, t1 = timed(step_device, args.steps, record_sizes=True)
    prof = eng.profile_report()
    eng.set_profiling(False)
    launches = eng.launch_count - launches0
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_per_step = ms_total / args.steps
    value = total_job / (ms_per_step * 1e-3)

    # ---- end-to-end arm (host buffers through the C ABI) -----------------------------------------
    for _ in range(0 if args.skip_e2e else 2):
        step_e2e()
    barrier()
    walls = []
    for _ in range(1 if args.skip_e2e else args.steps):
        barrier()
        tw = time.perf_counter()
        stats = step_e2e()
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - tw)
    e2e_s = float(numpy.mean(walls))
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    # whole job, per step: every block size uploads the images and its block list, and reads predictions + PSNRs back
    n_images_job = sum(len(sd) for sd in seeds_of)
    h2d = sum(n_images_job * HEIGHT * WIDTH_IMAGE + 12 * blocks[w] for w, _ in WIDTHS)
    d2h = sum(blocks[w] * (w * w + 8) for w, _ in WIDTHS)

    if rank == 0:
        per_size = {}
        for w, _ in WIDTHS:
            ms_w = float(numpy.mean([a.elapsed_time(b) for a, b in size_events[w]]))
            per_size[str(w)] = {'predictions_per_s_per_gpu': per_w[w]['n'] / (ms_w * 1e-3), 'ms': ms_w,
                                'tflops_algorithmic': 2. * MACS[w] * per_w[w]['n'] / (ms_w * 1e-3) / 1e12}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = peaks.get('bf16_tflops_sustained', 1590.0 if not peaks else None) or 1590.0
        achieved = prof['gemm_flops'] / (prof['gemm_ms'] * 1e-3) / 1e12 if prof['gemm_ms'] > 0 else 0.
        # DRAM bytes of one launch of the dominant kernel: read from the committed summary of the `ncu --set full` capture
        # (profiles/roofline_traffic.json, written by tools/ncu_summaries.py), not a literal
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')))
        except Exception:
            pass
        roofline = {
            'bound': 'tensor', 'kernel': 'gemm_tc_kernel (tcgen05 bf16x3 implicit GEMM)', 'achieved': achieved, 'peak': peak,
            'unit': 'TFLOP/s', 'frac': achieved / peak,
            'traffic': traffic.get('dram_bytes_per_launch'), 'traffic_launch': traffic.get('launch'),
            'traffic_algorithmic_bytes': traffic.get('algorithmic_bytes_per_launch'), 'traffic_source': traffic.get('source'),
            'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)' if peaks
                           else 'fallback 1.59 PFLOP/s (B200_PROFILING.md)',
            'mma_passes': 3, 'frac_of_tensor_issue': 3. * achieved / peak,
            'gemm_launches_per_step': prof['gemm_launches'] / args.steps,
            'gemm_ms_per_step': prof['gemm_ms'] / args.steps, 'other_kernels_ms_per_step': prof['other_ms'] / args.steps,
            'note': 'achieved = algorithmic FLOPs (2*M*N*K of every GEMM-shaped layer launch) / summed CUDA-event time of those '
                    'launches in the timed region; each algorithmic MAC costs 3 bf16 MMA passes (hi*hi + hi*lo + lo*hi)',
        }
        line = {
            'metric': METRIC, 'value': value, 'unit': 'predictions/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': None, 'dtype': 'bf16x3 (bf16 hi/lo operands, 3 MMA passes, f32 accumulate)'
            if args.precision == 'bf16x3' else 'f32', 'data': 'synthetic', 'config': cfg,
            'per_size': per_size, 'roofline': roofline,
            'e2e': {'value': total_job / e2e_s, 'unit': 'predictions/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': 1e3 * e2e_s,
                    'mean_psnr_pnn': stats['mean_psnr_pnn'], 'frequency_win_pnn': stats['frequency_win_pnn']},
            'gpu_launches': launches, 'clocks': clocks,
            'hevc_baseline': {'what': 'best of the 35 HEVC intra modes per block (psnrs_hevc_best_mode), computed once outside the '
                                      'timed region and supplied as the baseline PSNR array', 'ms': hevc_baseline_ms,
                              'blocks_per_s': total / (hevc_baseline_ms * 1e-3), 'mean_psnr': float(base_host.mean())},
        }
        if not args.no_cpu_baseline and world == 1:
            t_cpu = time.perf_counter()
            rates, counts = cpu_reference_rates(tensors, args.cpu_baseline_seconds)
            line['cpu_baseline'] = {
                'value': aggregate_rate(rates, blocks), 'unit': 'predictions/s', 'cores': os.cpu_count() or 1, 'kind': 'port',
                'sample': 'network forward only (contexts pre-gathered), batches of 10 as pnn/batching.py drives sess.run, '
                          '%s blocks per size, extrapolated to the full block mix' % {str(k): v for k, v in counts.items()},
                'per_size': {str(w): rates[w] for w in rates}, 'seconds': time.perf_counter() - t_cpu,
                'note': 'fp32 oracle on torch-CPU (oneDNN/MKL) standing in for the reference TF-CPU path (TensorFlow not installable)',
            }
        if args.report:
            sys.stderr.write(prof['text'] + '\n')
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == '__main__':
    main()
