"""Developer aid (GPU box): error table of every net / precision against the CPU oracle."""
import os
import sys
import tempfile

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import helpers                                                                    # noqa: E402
from context_adaptive_neural_network_based_prediction_b200 import Engine          # noqa: E402


def main():
    eng = Engine()
    tmp = tempfile.mkdtemp()
    cases = [(4, True, 1.6, 1500), (8, True, 1.6, 400), (4, False, 2.0, 300), (8, False, 2.0, 300), (16, False, 1.6, 200),
             (32, False, 1.5, 40), (64, False, 1.4, 4)]
    if len(sys.argv) > 1:
        want = set(sys.argv[1].split(','))
        cases = [c for c in cases if '%s%d' % ('fc' if c[1] else 'conv', c[0]) in want]
    for width, is_fc, gain, limit in cases:
        path, wts = helpers.make_net_file(tmp, width, is_fc, seed=width, gain=gain)
        eng.load_net(path)
        images = numpy.stack([helpers.synthetic_image(max(96, 3 * width), max(128, 4 * width), s) for s in range(2)])
        rows, cols, idx = [], [], []
        for i in range(2):
            r, c = helpers.grid_blocks(images.shape[1], images.shape[2], width)
            rows.append(r); cols.append(c); idx.append(numpy.full(len(r), i, dtype=numpy.int32))
        rows, cols, idx = numpy.concatenate(rows)[:limit], numpy.concatenate(cols)[:limit], numpy.concatenate(idx)[:limit]
        pred, u8, psnrs, _ = helpers.oracle_predict_blocks(wts, width, is_fc, images, idx, rows, cols)
        res = {}
        for prec in ('fp32', 'bf16x3'):
            eng.set_precision(prec)
            out = eng.predict_image_blocks(width, is_fc, images, rows, cols, idx)
            res[prec] = out
            e = numpy.abs(out['predictions_float32'] - pred)
            bad = numpy.argwhere(e > 1e-2)
            print('%s-%d %-6s n=%d |out|max=%.2f  err max %.3e mean %.3e  same-u8 %.5f  bad blocks %s'
                  % ('FC' if is_fc else 'CONV', width, prec, len(rows), numpy.abs(pred).max(), e.max(), e.mean(),
                     (out['predictions_uint8'] == u8).mean(), sorted(set(bad[:, 0].tolist()))[:12]), flush=True)
        e = numpy.abs(res['fp32']['predictions_float32'] - res['bf16x3']['predictions_float32'])
        print('      fp32 vs bf16x3: max %.3e' % e.max(), flush=True)


if __name__ == '__main__':
    main()
