"""One in-loop call per convolutional width (after warm-up), for `ncu --metrics gpu__time_duration.sum`: which kernels a call is made of."""
import os, sys, tempfile
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers
from context_adaptive_neural_network_based_prediction_b200 import Engine
eng = Engine(); tmp = tempfile.mkdtemp()
eng.set_hm_fused('--no-fused' not in sys.argv)
widths = [int(w) for w in sys.argv[1].split(',')]
for width in widths:
    path, _ = helpers.make_net_file(tmp, width, False, seed=width)
    eng.load_net(path)
    plane = helpers.synthetic_image(3 * width + 8, 3 * width + 24, 5).astype(numpy.int32)
    units = 2 * width // 4
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    for _ in range(3):
        eng.set_context(width, plane, width + 3, width + 5, flags, int(flags.sum()))
        eng.predict_hm(width)
eng.close()
