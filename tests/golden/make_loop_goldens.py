"""float64 goldens of the two pretrained checkpoints the reference ships (CONV-4, CONV-8), from the literal loop
implementation of oracle/nets.py (forward_conv_loops_float64: no library convolution, float64 throughout) on real image
blocks.  An independent pin of the torch oracle AND of the CUDA path: tests/test_oracle_cpu.py compares the torch oracle,
tests/test_gpu_parity.py the library, against tests/golden/loop_real.npz.  Regenerate:  python tests/golden/make_loop_goldens.py
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from context_adaptive_neural_network_based_prediction_b200 import weights as W   # noqa: E402
from oracle import context, nets                                                   # noqa: E402

MEAN = 117.8952234192841


def main():
    img = numpy.load(os.path.join(HERE, 'cliff_luma.npy'))
    rng = numpy.random.default_rng(11)
    out = {}
    for width in (4, 8):
        _, _, wts = W.load_flat(os.path.join(HERE, 'conv%d_single.pnnw' % width))
        n = 240
        rows = rng.integers(width, img.shape[0] - 2 * width + 1, n).astype(numpy.int32)
        cols = rng.integers(width, img.shape[1] - 2 * width + 1, n).astype(numpy.int32)
        for masks in ((0, 0), (4, 4)):
            above, left, flat, targets = context.gather_image_blocks(img[None], numpy.zeros(n, int), rows, cols, width, MEAN,
                                                                     masks[0], masks[1])
            pred = nets.forward_conv_loops_float64(wts, above, left)[..., 0]
            out['pred_%d_m%d%d' % (width, masks[0], masks[1])] = pred
        out['rows_%d' % width], out['cols_%d' % width] = rows, cols
    numpy.savez_compressed(os.path.join(HERE, 'loop_real.npz'), **out)
    print('wrote loop_real.npz')


if __name__ == '__main__':
    main()
