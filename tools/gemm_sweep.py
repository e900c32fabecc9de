"""Tuning aid (GPU box): TFLOP/s of the tcgen05 GEMM kernel over problem shapes and with loads / stores disabled."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from context_adaptive_neural_network_based_prediction_b200 import Engine
eng = Engine()
M = 262144
print('%8s %6s %6s %6s  %8s %8s' % ('M', 'N', 'K', 'flags', 'ms', 'TFLOP/s'))
shapes = [(M, 1200, 1200), (M, 1280, 1200), (M, 1200, 1216), (M, 1200, 1280), (M, 1280, 1280), (M, 256, 1200), (M, 256, 1280),
          (M, 256, 2304), (M, 512, 1280), (M, 128, 1152), (M, 64, 576), (4 * M, 64, 576), (M, 1200, 80), (M, 1200, 320), (M, 16, 1200)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]]
for (m, n, k) in shapes:
    for flags in [int(v) for v in os.environ.get("FLAGS", "0,1,2,4,3,7").split(",")]:
        ms = eng.time_gemm(m, n, k, 5, flags)
        print('%8d %6d %6d %6d  %8.3f %8.1f' % (m, n, k, flags, ms, 2. * m * n * k / ms / 1e9), flush=True)
