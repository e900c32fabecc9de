"""Conv in-loop calls in a tight loop while nvidia-smi samples the SM clock: are these latency-chain kernels running at idle clocks?"""
import os, subprocess, sys, tempfile, threading, time
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers
from context_adaptive_neural_network_based_prediction_b200 import Engine
eng = Engine(); tmp = tempfile.mkdtemp()
lines = []
proc = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,pstate', '--format=csv,noheader', '-lms', '100'], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()
for width in (16, 64):
    path, _ = helpers.make_net_file(tmp, width, False, seed=width)
    eng.load_net(path)
    plane = helpers.synthetic_image(3 * width + 8, 3 * width + 24, 5).astype(numpy.int32)
    units = 2 * width // 4
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    t0 = time.time(); n = 0; dev = []
    while time.time() - t0 < 2.5:
        eng.set_context(width, plane, width + 3, width + 5, flags, int(flags.sum()))
        eng.predict_hm(width); n += 1; dev.append(eng.last_hm_device_ms)
    t1 = time.time()
    clk = [l for t, l in lines if t0 + 0.5 < t < t1]
    print('W=%d: %.1f us wall per call, device median %.1f us (first 50: %.1f, last 50: %.1f); clocks during the loop: %s' % (
        width, (t1 - t0) / n * 1e6, numpy.median(dev) * 1e3, numpy.median(dev[:50]) * 1e3, numpy.median(dev[-50:]) * 1e3, sorted(set(clk))[:6]))
proc.terminate()
eng.close()
