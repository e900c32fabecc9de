"""Dump true TensorFlow outputs of a frozen PNN graph as a golden fixture ("TF parity mode").

RUN THIS WHERE TENSORFLOW 1.x IS INSTALLED (it is not in this repository's image, which is why the parity of the
network arithmetic against TensorFlow is still unpinned).  It feeds seeded contexts to the frozen
graph through the node names the HM integration uses (reference TComPrediction.cpp:570-571, 590-599) and writes

    tests/golden/tf_<fc|conv><W>.npz : inputs, `predictions` (the raw node_output), width, is_fc
    tests/golden/tf_<fc|conv><W>.pnnw: the same weights as a PNNW flat binary (no TensorFlow needed for this part)

`tests/test_tf_goldens.py` picks every such pair up: the CPU test checks the oracle against `predictions`, the GPU
test checks libpnn_cuda, both to the 1e-2 / 99.9 % bar of BASELINE.json.

    python tools/dump_tf_goldens.py --frozen-graph .../graph_output.pbtxt --width 8 --fc [--n 64] [--seed 0]
"""
import argparse
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from context_adaptive_neural_network_based_prediction_b200 import weights as W   # noqa: E402

MEAN = 117.8952234192841


def seeded_contexts(width, n, seed):
    """Mean-centred contexts of smooth random uint8 content, some with masked tails (zeros), as the reference feeds."""
    rng = numpy.random.RandomState(seed)
    above = numpy.zeros((n, width, 3 * width, 1), dtype=numpy.float32)
    left = numpy.zeros((n, 2 * width, width, 1), dtype=numpy.float32)
    for i in range(n):
        base, amp = rng.uniform(40, 200), rng.uniform(5, 60)
        fy, fx = rng.uniform(0.02, 0.5, size=2)
        yy, xx = numpy.mgrid[0:3 * width, 0:3 * width]
        img = numpy.clip(base + amp * numpy.sin(fy * yy + fx * xx) + rng.normal(0, 3, yy.shape), 0, 255).round()
        above[i, :, :, 0] = img[:width, :] - MEAN
        left[i, :, :, 0] = img[width:, :width] - MEAN
        if i % 4 == 1:
            above[i, :, 2 * width:, 0] = 0.            # masked above-right
        if i % 4 == 2:
            left[i, width:, :, 0] = 0.                 # masked below-left
    return above, left


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frozen-graph', required=True)
    ap.add_argument('--width', type=int, required=True, choices=(4, 8, 16, 32, 64))
    kind = ap.add_mutually_exclusive_group(required=True)
    kind.add_argument('--fc', action='store_true')
    kind.add_argument('--conv', action='store_true')
    ap.add_argument('--n', type=int, default=64)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--out-dir', default=os.path.join(ROOT, 'tests', 'golden'))
    a = ap.parse_args()
    import tensorflow as tf
    if hasattr(tf, 'compat') and hasattr(tf.compat, 'v1') and not hasattr(tf, 'Session'):
        tf = tf.compat.v1
        tf.disable_eager_execution()

    w = a.width
    above, left = seeded_contexts(w, a.n, a.seed)
    graph_def = tf.GraphDef()
    with open(a.frozen_graph, 'rb') as f:
        graph_def.ParseFromString(f.read())
    # the frozen graphs have batch size 1 (freezing_graph_pnn.py:100-102): one run per context, as HM does
    with tf.Graph().as_default() as graph:
        tf.import_graph_def(graph_def, name='')
        with tf.Session(graph=graph) as sess:
            if a.fc:
                # flattened context = above (row-major) then left (reference batching / extraction order)
                flat = numpy.concatenate([above.reshape(a.n, -1), left.reshape(a.n, -1)], axis=1)
                out_name = 'fully_connected/node_output:0'
                preds = [sess.run(out_name, {'node_flattened_context:0': flat[i:i + 1]}) for i in range(a.n)]
            else:
                n_t = len(W.STRIDES_BRANCH[w])
                out_name = 'convolutional/merger/transpose_convolution_%d/node_output:0' % (n_t - 1)
                preds = [sess.run(out_name, {'node_portion_above:0': above[i:i + 1], 'node_portion_left:0': left[i:i + 1]})
                         for i in range(a.n)]
    predictions = numpy.concatenate(preds, axis=0).reshape(a.n, w, w).astype(numpy.float32)
    stem = os.path.join(a.out_dir, 'tf_%s%d' % ('fc' if a.fc else 'conv', w))
    W.export_frozen_graph(a.frozen_graph, w, a.fc, stem + '.pnnw')
    numpy.savez_compressed(stem + '.npz', above=above, left=left, predictions=predictions, width=w, is_fc=int(a.fc),
                           tensorflow_version=str(getattr(tf, '__version__', getattr(tf, 'VERSION', '?'))))
    print('wrote %s.npz and %s.pnnw' % (stem, stem))


if __name__ == '__main__':
    main()
