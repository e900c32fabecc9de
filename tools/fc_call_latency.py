import os, sys, time, ctypes, tempfile, numpy
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import helpers
from context_adaptive_neural_network_based_prediction_b200 import Engine
eng = Engine(); tmp = tempfile.mkdtemp()
for width in (4, 8):
    path, _ = helpers.make_net_file(tmp, width, True, seed=width); eng.load_net(path)
    ctx = numpy.random.default_rng(0).normal(0, 30, 5 * width * width).astype(numpy.float32)
    out = numpy.zeros(width * width, dtype=numpy.float32)
    lib, h = eng._lib, eng._h
    lib.pnn_predict_hm_context.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    f = lib.pnn_predict_hm_context
    a, o = ctx.ctypes.data, out.ctypes.data
    for _ in range(200): f(h, width, a, None, o)
    n = 3000
    t0 = time.perf_counter()
    for _ in range(n): f(h, width, a, None, o)
    t1 = time.perf_counter()
    # python call overhead of the same ctypes signature on a no-op: pnn_launch_count
    g = lib.pnn_launch_count; g.argtypes = [ctypes.c_void_p]
    t2 = time.perf_counter()
    for _ in range(n): g(h)
    t3 = time.perf_counter()
    print('width %d: %.2f us per call (ctypes no-op %.2f us)' % (width, (t1 - t0) / n * 1e6, (t3 - t2) / n * 1e6))
