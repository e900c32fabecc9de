"""Developer aid: layer-by-layer comparison of FC-8 with a float64 forward."""
import os, sys, tempfile
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers
from context_adaptive_neural_network_based_prediction_b200 import Engine
from oracle import context
eng = Engine(); tmp = tempfile.mkdtemp()
for width in (8,):
    path, wts = helpers.make_net_file(tmp, width, True, seed=width, gain=1.6)
    eng.load_net(path)
    images = numpy.stack([helpers.synthetic_image(96, 128, s) for s in range(2)])
    r, c = helpers.grid_blocks(96, 128, width)
    idx = numpy.zeros(len(r), dtype=numpy.int32)
    above, left, flat, _ = context.gather_image_blocks(images, idx, r, c, width, helpers.MEAN, 0, 0)
    acts = [flat.astype(numpy.float64)]
    for i in range(4):
        x = acts[-1] @ wts['fully_connected/weights_%d' % i].astype(numpy.float64) + wts['fully_connected/biases_%d' % i]
        if i != 3: x = numpy.maximum(0.1 * x, x)
        acts.append(x)
    for prec in ('fp32', 'bf16x3'):
        eng.set_precision(prec)
        g = eng.predict_batch(width, True, flat).reshape(len(r), -1)
        for b in range(4):
            a = eng.get_activation(width, True, b, len(r))
            e = numpy.abs(a - acts[b])
            w = numpy.unravel_index(numpy.argmax(e), e.shape)
            cols = numpy.unique(numpy.argwhere(e > 50 * numpy.median(e) + 1e-7)[:, 1])
            print('%s buffer %d: max err %.3e at %s (ref %.4f got %.4f) |ref|max %.2f; outlier cols %s'
                  % (prec, b, e.max(), w, acts[b][w], a[w], numpy.abs(acts[b]).max(), cols[:20]))
        e = numpy.abs(g - acts[4]); w = numpy.unravel_index(numpy.argmax(e), e.shape)
        print('%s output: max err %.3e at %s' % (prec, e.max(), w))
