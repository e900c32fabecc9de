"""Golden vectors of the offline gather from the REFERENCE ITSELF (run in the build container, where /root/reference exists).

Imports the reference's sets/common.py unmodified (it only needs numpy) and runs
`extract_context_portions_targets_from_channels_plus_preprocessing` (sets/common.py:265-349: slicing :13-110, mean
subtraction and masks :351-475, FC flattening :467-472) on small crops; the outputs are committed as
tests/golden/reference_gather.npz and pin oracle.context.gather_image_blocks (tests/test_oracle_cpu.py) and, through it,
the CUDA gather.  Regenerate with:  python tests/golden/make_reference_gather_golden.py
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
import sets.common as ref_common          # noqa: E402

MEAN = 117.8952234192841


def main():
    luma = numpy.load(os.path.join(HERE, 'cliff_luma.npy'))
    rng = numpy.random.default_rng(7)
    images = numpy.stack([luma[20:20 + 72, 30:30 + 104], luma[80:80 + 72, 120:120 + 104],
                          rng.integers(0, 256, (72, 104)).astype(numpy.uint8)])
    out = {'images': images}
    for width in (4, 8, 16):
        n_pos = 5
        row_1sts = rng.integers(0, images.shape[1] - 3 * width + 1, n_pos)
        col_1sts = rng.integers(0, images.shape[2] - 3 * width + 1, n_pos)
        out['rows_%d' % width] = row_1sts
        out['cols_%d' % width] = col_1sts
        for masks in ((0, 0), (4, 0), (0, width), (width, 4)):
            for is_fc in (True, False):
                res = ref_common.extract_context_portions_targets_from_channels_plus_preprocessing(
                    images[..., None], width, row_1sts, col_1sts, MEAN, masks, is_fc)
                tag = '%d_%d_%d_%s' % (width, masks[0], masks[1], 'fc' if is_fc else 'conv')
                if is_fc:
                    out['flat_' + tag], out['target_' + tag] = res
                else:
                    out['above_' + tag], out['left_' + tag], out['target_' + tag] = res
    numpy.savez_compressed(os.path.join(HERE, 'reference_gather.npz'), **out)
    print('wrote reference_gather.npz with %d arrays' % len(out))


if __name__ == '__main__':
    main()
