"""Generates the committed golden fixtures of tests/golden/ from the reference tree.

Run HERE (container with /root/reference), never on the GPU box:
    python tests/golden/make_golden.py

Fixtures:
  conv4_single.pnnw, conv8_single.pnnw
      the two pretrained checkpoints the reference ships with data
      (pnn/results/width_target_{4,8}/convolutional/single/luminance/1_0/masks_tr_random/model_800000.ckpt),
      exported to the PNNW flat binary by weights.export_checkpoint.
  cliff_luma.npy
      160 x 240 uint8 luminance crop of the reference's sets/pseudo_data/rgb_cliff.jpg
      (ITU-R BT.601 luma of the decoded JPEG).
  conv_real.npz
      block positions on that image and the fp32 CPU oracle's outputs for CONV-4 / CONV-8 with the real
      weights, masks (0, 0) and (4, 4).
  extract_ref.npz
      inputs and outputs of the reference's own extract_context_portions (compiled unmodified into
      oracle/_ref/libextract_ref.so) on random planes and random availability patterns.
  hevc_ref.npz
      intra patterns and the 35 predictions of the reference's own hevc_intraprediction (compiled unmodified into
      oracle/_ref/libhevc_intra_ref.so) for widths 4 .. 64, with and without masks.
"""
import ctypes
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from context_adaptive_neural_network_based_prediction_b200 import weights as W   # noqa: E402
from oracle import context, epilogue, nets                                        # noqa: E402

REF = '/root/reference'
MEAN = 117.8952234192841


def main():
    from PIL import Image
    rng = numpy.random.default_rng(2024)
    rgb = numpy.asarray(Image.open(os.path.join(REF, 'sets/pseudo_data/rgb_cliff.jpg')).convert('RGB')).astype(numpy.float64)
    luma = numpy.clip(numpy.round(0.299 * rgb[..., 0] + 0.587 * rgb[..., 1] + 0.114 * rgb[..., 2]), 0, 255).astype(numpy.uint8)
    crop = numpy.ascontiguousarray(luma[200:360, 300:540])
    numpy.save(os.path.join(HERE, 'cliff_luma.npy'), crop)

    out = {}
    for width in (4, 8):
        prefix = os.path.join(REF, 'pnn/results/width_target_%d/convolutional/single/luminance/1_0/masks_tr_random/model_800000.ckpt' % width)
        wts = W.export_checkpoint(prefix, width, False, os.path.join(HERE, 'conv%d_single.pnnw' % width))
        n = 256
        rows = rng.integers(width, crop.shape[0] - 2 * width + 1, n).astype(numpy.int32)
        cols = rng.integers(width, crop.shape[1] - 2 * width + 1, n).astype(numpy.int32)
        out['rows_%d' % width] = rows
        out['cols_%d' % width] = cols
        for masks in ((0, 0), (4, 4)):
            above, left, _, targets = context.gather_image_blocks(crop[None], numpy.zeros(n, int), rows, cols, width, MEAN, *masks)
            pred = nets.forward_conv(wts, above, left)[..., 0]
            u8 = epilogue.epilogue_numpy(pred, MEAN)
            psnrs = numpy.array([epilogue.psnr(targets[i], u8[i]) for i in range(n)])
            tag = '%d_m%d%d' % (width, masks[0], masks[1])
            out['pred_' + tag] = pred
            out['u8_' + tag] = u8
            out['psnr_' + tag] = psnrs
            print('CONV-%d masks %s: mean PSNR %.3f dB' % (width, masks, psnrs.mean()))
    numpy.savez_compressed(os.path.join(HERE, 'conv_real.npz'), **out)

    # reference extract_context_portions, compiled unmodified (oracle/Makefile target `ref`)
    lib = ctypes.CDLL(os.path.join(ROOT, 'oracle/_ref/libextract_ref.so'))
    fn = lib.ref_extract_context_portions
    fn.restype = ctypes.c_int
    cases = {}
    idx = 0
    for width in (4, 8, 16, 32, 64):
        for trial in range(6):
            stride = 3 * width + 40
            height = 3 * width + 8
            plane = rng.integers(0, 256, (height, stride)).astype(numpy.int32)
            units = 2 * width // 4
            flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
            if trial == 1:
                flags[:rng.integers(1, units // 2 + 1)] = 0                       # below-left missing
            elif trial == 2:
                flags[2 * units + 1 - rng.integers(1, units // 2 + 1):] = 0       # above-right missing
            elif trial == 3:
                flags[:units // 2] = 0
                flags[units + 1 + units // 2:] = 0
            elif trial == 4:
                flags[rng.integers(0, units)] = 0                                 # a hole on the left side
                flags[units + 1 + rng.integers(0, units)] = 0                     # a hole above
            elif trial == 5:
                flags[:units] = rng.integers(0, 2, units)
                flags[units + 1:] = rng.integers(0, 2, units)
            n_avail = int(flags.sum())
            mean = numpy.float32(MEAN if trial % 2 else 0.)
            orow, ocol = width + 2, width + 3
            above = numpy.full(3 * width * width, -7., dtype=numpy.float32)
            left = numpy.full(2 * width * width, -7., dtype=numpy.float32)
            fb = flags.astype(numpy.bool_)
            origin = plane.ctypes.data + 4 * (orow * stride + ocol)
            code = fn(ctypes.c_void_p(origin), above.ctypes.data_as(ctypes.c_void_p), left.ctypes.data_as(ctypes.c_void_p),
                      fb.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n_avail), 4, 4, units, units, width, width, stride,
                      ctypes.c_float(mean))
            assert code == 0
            p = 'c%d_' % idx
            cases[p + 'plane'] = plane.astype(numpy.uint8)   # values are 0..255; tests cast back to int32
            cases[p + 'flags'] = flags
            cases[p + 'meta'] = numpy.array([width, orow, ocol, n_avail], dtype=numpy.int32)
            cases[p + 'mean'] = numpy.array([mean], dtype=numpy.float32)
            cases[p + 'above'] = above
            cases[p + 'left'] = left
            idx += 1
    cases['n_cases'] = numpy.array([idx], dtype=numpy.int32)
    numpy.savez_compressed(os.path.join(HERE, 'extract_ref.npz'), **cases)
    print('wrote %d extraction cases' % idx)

    # reference hevc_intraprediction, compiled unmodified (oracle/Makefile target `ref`)
    lib = ctypes.CDLL(os.path.join(ROOT, 'oracle/_ref/libhevc_intra_ref.so'))
    hevc = {}
    idx = 0
    for width, n_patterns in ((4, 3), (8, 3), (16, 2), (32, 1), (64, 1)):
        for trial in range(n_patterns):
            mask_w = 0 if trial == 0 else 4 * int(rng.integers(0, width // 4 + 1))
            mask_h = 0 if trial == 0 else 4 * int(rng.integers(0, width // 4 + 1))
            hp, wp = 2 * width + 1 - mask_h, 2 * width + 1 - mask_w
            yy, xx = numpy.mgrid[0:hp, 0:wp]
            pattern = numpy.clip(110 + 3 * xx - 2 * yy + rng.integers(-12, 13, (hp, wp)), 0, 255).astype(numpy.uint8)
            preds = numpy.zeros((35, width, width), dtype=numpy.uint8)
            for mode in range(35):
                out = numpy.zeros(width * width, dtype=numpy.uint8)
                assert lib.ref_hevc_intraprediction(hp, wp, width, pattern.ctypes.data_as(ctypes.c_void_p),
                                                    out.ctypes.data_as(ctypes.c_void_p), mode) == 0
                preds[mode] = out.reshape(width, width)
            hevc['h%d_row' % idx] = pattern[0].copy()
            hevc['h%d_col' % idx] = pattern[:, 0].copy()
            hevc['h%d_width' % idx] = numpy.array([width], dtype=numpy.int32)
            hevc['h%d_preds' % idx] = preds
            idx += 1
    hevc['n_cases'] = numpy.array([idx], dtype=numpy.int32)
    numpy.savez_compressed(os.path.join(HERE, 'hevc_ref.npz'), **hevc)
    print('wrote %d HEVC intra cases' % idx)


if __name__ == '__main__':
    main()
