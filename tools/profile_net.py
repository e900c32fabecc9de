"""Profiling aid (GPU box): runs ONE net a few times on random contexts so that the launch order under ncu is known.

    ncu --clock-control none -k regex:gemm_tc -s <13 * warm> -c 13 ... python tools/profile_net.py 16 32768 3

CONV-16 GEMM launches per call, in order: above (64,576) (128,1600) (128,1152); left: the same three; tconv0 (128,1152);
tconv1 phases (64,512) (64,768) (64,768) (64,1152); tconv2 (64,576); last (32,64).
"""
import os
import sys
import tempfile

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from context_adaptive_neural_network_based_prediction_b200 import Engine, weights as W   # noqa: E402

width = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 3
is_fc = width <= 8 and os.environ.get('CONV', '0') == '0'
eng = Engine()
path = os.path.join(tempfile.mkdtemp(), 'net.pnnw')
W.save_flat(path, width, is_fc, W.init_weights(width, is_fc, seed=1, bias_std=0.05, gain=1.8))
eng.load_net(path)
rng = numpy.random.default_rng(0)
if is_fc:
    args = (rng.normal(0., 40., (n, 5 * width * width)).astype(numpy.float32),)
else:
    args = (rng.normal(0., 40., (n, width, 3 * width, 1)).astype(numpy.float32),
            rng.normal(0., 40., (n, 2 * width, width, 1)).astype(numpy.float32))
eng.set_profiling(True)
for _ in range(calls):
    out = eng.predict_batch(width, is_fc, *args)
print(eng.profile_report()['text'])
