"""Batch-1 latency of the CPU baseline backend (oracle/_ref/libpnn_ref.so, libtorch-CPU) per block width and per
intra-op thread count: picks the thread setting the HM baseline leg is run with (hm/run_hm.py --backend cpu).
Baseline infrastructure, not product."""
import ctypes, json, os, sys, tempfile, time
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers
lib = ctypes.CDLL(os.path.join(ROOT, 'oracle', '_ref', 'libpnn_ref.so'))
h = ctypes.c_void_p()
lib.pnn_create.argtypes = [ctypes.c_char_p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
assert lib.pnn_create(None, 0., 1, 0, ctypes.byref(h)) == 0
lib.pnn_load_net.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
lib.pnn_predict_hm_context.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
lib.pnn_ref_set_threads.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
tmp = tempfile.mkdtemp()
cores = os.cpu_count()
counts = sorted(set(c for c in (1, 2, 4, 8, 16, 32, cores) if c <= cores))
best = {}
print('%5s' % 'W' + ''.join('%10d' % c for c in counts) + '   (us per call, threads across)')
for w in (4, 8, 16, 32, 64):
    fc = w <= 8
    path, _ = helpers.make_net_file(tmp, w, fc, seed=w)
    assert lib.pnn_load_net(h, path.encode()) == 0
    rng = numpy.random.default_rng(w)
    a = rng.normal(0, 40, 3 * w * w).astype(numpy.float32)
    l = rng.normal(0, 40, 2 * w * w).astype(numpy.float32)
    flat = numpy.concatenate([a, l])
    out = numpy.zeros(w * w, numpy.float32)
    row = []
    for c in counts:
        lib.pnn_ref_set_threads(h, c, c)
        args = (h, w, flat.ctypes.data, None, out.ctypes.data) if fc else (h, w, a.ctypes.data, l.ctypes.data, out.ctypes.data)
        n = 300 if fc else (60 if w < 64 else 25)
        for _ in range(5):
            lib.pnn_predict_hm_context(*args)
        t0 = time.perf_counter()
        for _ in range(n):
            lib.pnn_predict_hm_context(*args)
        row.append((time.perf_counter() - t0) / n * 1e6)
    best[w] = (counts[int(numpy.argmin(row))], min(row))
    print('%5d' % w + ''.join('%10.1f' % r for r in row))
print(json.dumps({'host_cores': cores, 'best_threads_and_us': {str(w): best[w] for w in best}}))
