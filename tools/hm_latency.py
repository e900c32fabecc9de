"""In-loop (HM) call pair latency per block size: wall time of pnn_set_context + pnn_predict_hm, device time, CPU oracle beside it."""
import os, sys, tempfile, time
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers
from context_adaptive_neural_network_based_prediction_b200 import Engine, weights
cpu = '--cpu' in sys.argv
fused = '--no-fused' not in sys.argv
eng = Engine(); tmp = tempfile.mkdtemp()
eng.set_hm_fused(fused)
if "--prof" in sys.argv: eng.set_profiling(True)
params = {4: 2998816, 8: 3344464, 16: 1339073, 32: 5622657, 64: 20652545}
print('%3s %10s %10s %12s %12s' % ('W', 'wall_us', 'device_us', 'GB/s(params)', 'cpu_oracle_us'))
results = {}
for width in (4, 8, 16, 32, 64):
    is_fc = width <= 8
    path, wts = helpers.make_net_file(tmp, width, is_fc, seed=width)
    eng.load_net(path)
    plane = helpers.synthetic_image(3 * width + 8, 3 * width + 24, 5).astype(numpy.int32)
    units = 2 * width // 4
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    for _ in range(5):
        eng.set_context(width, plane, width + 3, width + 5, flags, int(flags.sum()))
        eng.predict_hm(width)
    n = 200 if width < 64 else 100
    dev = []
    t0 = time.perf_counter()
    for _ in range(n):
        eng.set_context(width, plane, width + 3, width + 5, flags, int(flags.sum()))
        eng.predict_hm(width)
        dev.append(eng.last_hm_device_ms)
    wall = (time.perf_counter() - t0) / n * 1e6
    cpu_us = float('nan')
    if cpu:
        import torch
        from oracle import context, nets
        torch.set_num_threads(os.cpu_count())
        code, above, left = context.extract_context_portions_hm(plane.ravel(), plane.shape[1], (width + 3) * plane.shape[1] + width + 5,
                                                                flags, int(flags.sum()), 4, 4, units, units, width, helpers.MEAN)
        args = (numpy.concatenate([above, left])[None],) if is_fc else (above.reshape(1, width, 3 * width, 1), left.reshape(1, 2 * width, width, 1))
        nets.forward(wts, width, is_fc, args)
        t0 = time.perf_counter()
        for _ in range(20):
            nets.forward(wts, width, is_fc, args)
        cpu_us = (time.perf_counter() - t0) / 20 * 1e6
    # the persistent FC kernel serves a call without a launch: there are no events to time it with, only the wall clock
    d = float(numpy.median(dev)) * 1e3
    print('%3d %10.1f %10s %12.1f %12.1f' % (width, wall, '%.1f' % d if d > 0 else 'n/a (persistent)',
                                             params[width] * 4 / ((d if d > 0 else wall) * 1e-6) / 1e9, cpu_us), flush=True)
    results[width] = (wall, cpu_us)
