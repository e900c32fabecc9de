"""Small workload touching every kernel (offline FC + conv, in-loop FC + conv) for compute-sanitizer runs."""
import os, sys, tempfile
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers
from context_adaptive_neural_network_based_prediction_b200 import Engine
eng = Engine(); tmp = tempfile.mkdtemp()
img = numpy.stack([helpers.synthetic_image(96, 128, s) for s in range(2)])
for width, is_fc in ((4, True), (8, False), (16, False), (32, False)):
    path, _ = helpers.make_net_file(tmp, width, is_fc, seed=width)
    eng.load_net(path)
    r, c = helpers.grid_blocks(96, 128, width)
    idx = numpy.zeros(len(r), dtype=numpy.int32); idx[len(r) // 2:] = 1
    for prec in ('bf16x3', 'fp32'):
        eng.set_precision(prec)
        out = eng.predict_image_blocks(width, is_fc, img, r[:150], c[:150], idx[:150])
    eng.set_precision('bf16x3')
    plane = helpers.synthetic_image(3 * width + 8, 3 * width + 24, 5).astype(numpy.int32)
    units = 2 * width // 4
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8); flags[:units // 2] = 0
    if width != 8:
        for fused in (True, False):
            eng.set_hm_fused(fused)
            eng.set_context(width, plane, width + 3, width + 5, flags, int(flags.sum()))
            eng.predict_hm(width)
        eng.set_hm_fused(True)
        eng.set_context(width, plane, width + 3, width + 5, flags, int(flags.sum()))
        eng.predict_hm_begin(width)                      # posted ahead, collected by the call that follows
        eng.predict_hm(width)
print('sanitize workload done')
