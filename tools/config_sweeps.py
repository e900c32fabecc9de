"""Measurements for BASELINE.json configs[0] (FC-8 over a Kodak-shaped image) and configs[2] (CONV-64 batch sweep).

Writes one JSON object per config to stdout; run on the GPU box:  python tools/config_sweeps.py > gpurun_out/configs.json
"""
import json, os, sys, tempfile, time
import numpy, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from context_adaptive_neural_network_based_prediction_b200 import Engine, offline, weights

MEAN = bench.MEAN
eng = Engine(mean_training=MEAN)
tmp = tempfile.mkdtemp()
dev = torch.device('cuda', 0)


def timed(fn, iters):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ---------------- configs[0]: FC-8 over all 8x8 blocks of a synthetic 512x768 (Kodak-shaped) luminance image
wts = weights.init_weights(8, True, seed=8)
path = os.path.join(tmp, 'fc8.pnnw'); weights.save_flat(path, 8, True, wts); eng.load_net(path)
img = bench.synthetic_image(512, 768, 0)
rows, cols = offline.grid_blocks(512, 768, 8)
n = len(rows)
for _ in range(3):
    eng.predict_image_blocks(8, True, img, rows, cols)
t0 = time.perf_counter()
for _ in range(20):
    out = eng.predict_image_blocks(8, True, img, rows, cols)
single_ms = (time.perf_counter() - t0) / 20 * 1e3
n_img = 1024
d_img = torch.from_numpy(img).to(dev).unsqueeze(0).repeat(n_img, 1, 1).contiguous()
idx, rr, cc = offline.blocks_of_images(n_img, 512, 768, 8)
d_idx, d_rows, d_cols = (torch.from_numpy(a).to(dev) for a in (idx, rr, cc))
total = len(rr)
d_u8 = torch.empty((total, 64), dtype=torch.uint8, device=dev)
d_psnr = torch.empty(total, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
ms = timed(lambda: eng.predict_image_blocks_device(8, True, d_img.data_ptr(), n_img, 512, 768, d_idx.data_ptr(), d_rows.data_ptr(),
                                                   d_cols.data_ptr(), total, (0, 0), None, d_u8.data_ptr(), d_psnr.data_ptr(), stream), 3)
print(json.dumps({'config': 'configs[0]: FC-8, all 8x8 blocks of a synthetic 768x512 Kodak-shaped image', 'blocks_per_image': n,
                  'single_image_host_api_ms': single_ms, 'single_image_predictions_per_s': n / (single_ms * 1e-3),
                  'steady_state_images': n_img, 'steady_state_ms': ms, 'steady_state_predictions_per_s': total / (ms * 1e-3),
                  'steady_state_tflops_algorithmic': 2 * 3340800 * total / (ms * 1e-3) / 1e12,
                  'mean_psnr_single_image': float(out['psnrs'].mean())}))
del d_img, d_u8, d_psnr, d_idx, d_rows, d_cols
torch.cuda.empty_cache()

# ---------------- configs[2]: CONV-64, random init, batch sweep
wts = weights.init_weights(64, False, seed=64)
path = os.path.join(tmp, 'conv64.pnnw'); weights.save_flat(path, 64, False, wts); eng.load_net(path)
pool = 1 << 15                                           # distinct contexts kept in HBM (2.7 GB); larger batches loop over them
g = torch.Generator(device=dev); g.manual_seed(64)
above = torch.clamp(torch.randn((pool, 64, 192, 1), generator=g, device=dev) * 40., -118., 137.)
left = torch.clamp(torch.randn((pool, 128, 64, 1), generator=g, device=dev) * 40., -118., 137.)
outb = torch.empty((pool, 64, 64, 1), dtype=torch.float32, device=dev)
sweep = []
for e in range(0, 21):
    b = 1 << e
    chunk = min(b, pool)
    reps = b // chunk
    def run():
        for _ in range(reps):
            eng.predict_batch_device(64, False, above.data_ptr(), left.data_ptr(), chunk, outb.data_ptr(), stream)
    iters = 20 if b <= 64 else (5 if b <= 4096 else 1)
    ms = timed(run, iters)
    sweep.append({'batch': b, 'ms': ms, 'predictions_per_s': b / (ms * 1e-3), 'tflops_algorithmic': 2 * 1180696576 * b / (ms * 1e-3) / 1e12})
    sys.stderr.write('CONV-64 batch %8d  %10.3f ms  %12.1f pred/s  %7.1f TFLOP/s\n' % (b, ms, b / (ms * 1e-3), sweep[-1]['tflops_algorithmic']))
print(json.dumps({'config': 'configs[2]: CONV-64, seeded random init, contexts N(0,40^2) clipped to [-118,137], batch sweep '
                            '(batches above %d loop over the same %d resident contexts)' % (pool, pool), 'sweep': sweep}))
