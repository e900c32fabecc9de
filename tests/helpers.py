"""Shared helpers of the test-suite (synthetic images, weights on disk, oracle wrappers)."""
import os

import numpy

from context_adaptive_neural_network_based_prediction_b200 import weights as W

MEAN = 117.8952234192841

# weight gains that bring the outputs of the randomly initialised nets to tens of pixel units (the trained
# nets' range); the reference initialisers alone give |output| < 1 for the deep convolutional nets
GAIN = {(4, True): 1.6, (8, True): 1.6, (4, False): 2.6, (8, False): 2.0, (16, False): 1.9, (32, False): 2.0, (64, False): 2.05}


def synthetic_image(height, width, seed):
    """clip(128 + 60 sin(x/17) + 40 cos(y/11) + N(0, 4^2))."""
    rng = numpy.random.default_rng(seed)
    y, x = numpy.mgrid[0:height, 0:width]
    img = 128. + 60. * numpy.sin(x / 17.) + 40. * numpy.cos(y / 11.) + rng.normal(0., 4., (height, width))
    return numpy.clip(numpy.round(img), 0, 255).astype(numpy.uint8)


def grid_blocks(height, width, w):
    """All W-aligned target blocks whose context anchor (r - W, c - W) lies inside the image."""
    rows, cols = numpy.meshgrid(numpy.arange(w, height - w + 1, w), numpy.arange(w, width - w + 1, w), indexing='ij')
    return rows.ravel().astype(numpy.int32), cols.ravel().astype(numpy.int32)


def make_net_file(directory, width, is_fc, seed, bias_std=0.05, gain=1.):
    wts = W.init_weights(width, is_fc, seed, bias_std=bias_std, gain=gain)
    path = os.path.join(directory, 'net_%d_%d_%d.pnnw' % (width, int(is_fc), seed))
    W.save_flat(path, width, is_fc, wts)
    return path, wts


def oracle_predict_blocks(wts, width, is_fc, images, idx, rows, cols, masks=(0, 0)):
    from oracle import context, epilogue, nets
    above, left, flat, targets = context.gather_image_blocks(images, idx, rows, cols, width, MEAN, masks[0], masks[1])
    pred = nets.forward(wts, width, is_fc, (flat,) if is_fc else (above, left))[..., 0]
    u8 = epilogue.epilogue_numpy(pred, MEAN)
    psnrs = numpy.array([epilogue.psnr(targets[i], u8[i]) for i in range(len(rows))])
    return pred, u8, psnrs, (above, left, flat, targets)


def check_parity(pred_gpu, pred_ref, u8_gpu=None, u8_ref=None, tol=1e-2, min_identical=0.999):
    """BASELINE.json north_star: max abs error <= 1e-2 pixel units, >= 99.9 % identical rounded pixels."""
    err = float(numpy.abs(pred_gpu.astype(numpy.float64) - pred_ref.astype(numpy.float64)).max())
    assert err <= tol, 'max abs error %.3e exceeds %.1e' % (err, tol)
    if u8_gpu is not None:
        same = float((u8_gpu == u8_ref).mean())
        assert same >= min_identical, 'only %.5f of the rounded pixels are identical' % same
        # a differing pixel can only be a rounding tie broken by an error below the tolerance
        diff = numpy.abs(u8_gpu.astype(int) - u8_ref.astype(int))
        assert diff.max() <= 1
    return err
