"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): max abs error <= 1e-2 pixel units on the raw float32 predictions and
>= 99.9 % identical rounded/clipped pixels; bit-exact for the integer work (gather, rounding of identical
floats, PSNR sums) and across repeated / re-batched calls.
"""
import os

import numpy
import pytest

import helpers
from helpers import MEAN

pytestmark = pytest.mark.gpu

TOL = {'bf16x3': 1e-2, 'fp32': 1e-3}


def _image_set(n, h, w):
    return numpy.stack([helpers.synthetic_image(h, w, s) for s in range(n)])


def _blocks(images, width, limit=None, seed=0):
    n_img, h, w = images.shape
    rows, cols, idx = [], [], []
    for i in range(n_img):
        r, c = helpers.grid_blocks(h, w, width)
        rows.append(r)
        cols.append(c)
        idx.append(numpy.full(len(r), i, dtype=numpy.int32))
    rows, cols, idx = numpy.concatenate(rows), numpy.concatenate(cols), numpy.concatenate(idx)
    if limit is not None and len(rows) > limit:
        sel = numpy.sort(numpy.random.default_rng(seed).choice(len(rows), limit, replace=False))
        rows, cols, idx = rows[sel], cols[sel], idx[sel]
    return rows, cols, idx


@pytest.mark.parametrize('precision', ['bf16x3', 'fp32'])
@pytest.mark.parametrize('width,is_fc,limit', [(4, True, 3000), (8, True, 2000), (16, False, 300), (32, False, 60),
                                              (64, False, 6), (4, False, 800), (8, False, 500)])
def test_image_blocks_parity(engine, weights_dir, width, is_fc, limit, precision):
    """Fused gather + net + epilogue + PSNR against the oracle; gain > 1 so that outputs span tens of pixel units."""
    path, wts = helpers.make_net_file(weights_dir, width, is_fc, seed=width + 100 * int(is_fc), gain=helpers.GAIN[(width, is_fc)])
    engine.load_net(path)
    engine.set_precision(precision)
    images = _image_set(2, max(96, 3 * width), max(128, 4 * width))
    rows, cols, idx = _blocks(images, width, limit)
    out = engine.predict_image_blocks(width, is_fc, images, rows, cols, idx)
    pred, u8, psnrs, _ = helpers.oracle_predict_blocks(wts, width, is_fc, images, idx, rows, cols)
    assert numpy.abs(pred).max() > 3., 'test inputs too weak to say anything'
    helpers.check_parity(out['predictions_float32'], pred, out['predictions_uint8'], u8, tol=TOL[precision])
    same = (out['predictions_uint8'] == u8).reshape(len(rows), -1).all(axis=1)
    numpy.testing.assert_allclose(out['psnrs'][same], psnrs[same], rtol=0, atol=1e-9)


@pytest.mark.parametrize('width', [4, 8])
def test_real_checkpoints_against_golden(engine, golden_dir, width):
    """The two pretrained checkpoints the reference ships, against the committed oracle outputs."""
    engine.load_net(os.path.join(golden_dir, 'conv%d_single.pnnw' % width))
    img = numpy.load(os.path.join(golden_dir, 'cliff_luma.npy'))
    gold = numpy.load(os.path.join(golden_dir, 'conv_real.npz'))
    rows, cols = gold['rows_%d' % width], gold['cols_%d' % width]
    for precision in ('bf16x3', 'fp32'):
        engine.set_precision(precision)
        for masks in ((0, 0), (4, 4)):
            out = engine.predict_image_blocks(width, False, img, rows, cols, masks=masks)
            tag = '%d_m%d%d' % (width, masks[0], masks[1])
            helpers.check_parity(out['predictions_float32'], gold['pred_' + tag], out['predictions_uint8'], gold['u8_' + tag],
                                 tol=TOL[precision])
            same = (out['predictions_uint8'] == gold['u8_' + tag]).reshape(len(rows), -1).all(axis=1)
            numpy.testing.assert_allclose(out['psnrs'][same], gold['psnr_' + tag][same], rtol=0, atol=1e-9)


@pytest.mark.parametrize('width,is_fc', [(8, True), (16, False)])
def test_masks_and_image_borders(engine, weights_dir, width, is_fc):
    """Masks in {0, 4, ..., W} (sets/common.py:444-461); context parts outside the image are masked like unavailable units."""
    path, wts = helpers.make_net_file(weights_dir, width, is_fc, seed=7, gain=helpers.GAIN[(width, is_fc)])
    engine.load_net(path)
    engine.set_precision('bf16x3')
    images = _image_set(1, 5 * width, 6 * width)
    rows, cols, idx = _blocks(images, width)       # includes the last block row / column: below-left / above-right outside
    for masks in ((0, 0), (4, 0), (0, width), (width, width)):
        out = engine.predict_image_blocks(width, is_fc, images, rows, cols, idx, masks=masks)
        pred, u8, _, _ = helpers.oracle_predict_blocks(wts, width, is_fc, images, idx, rows, cols, masks)
        helpers.check_parity(out['predictions_float32'], pred, out['predictions_uint8'], u8)
    from context_adaptive_neural_network_based_prediction_b200 import PnnError
    for bad in ((3, 0), (0, width + 4)):
        with pytest.raises(PnnError):
            engine.predict_image_blocks(width, is_fc, images, rows, cols, idx, masks=bad)
    with pytest.raises(PnnError):    # a target block leaving the image
        engine.predict_image_blocks(width, is_fc, images, numpy.array([5 * width - 1], dtype=numpy.int32),
                                    numpy.array([width], dtype=numpy.int32))


@pytest.mark.parametrize('width,is_fc', [(4, True), (8, True), (16, False)])
def test_predict_batch_matches_fused_gather(engine, weights_dir, width, is_fc):
    """pnn_predict_batch on pre-processed contexts == pnn_predict_image_blocks, bit for bit (same kernels, same order)."""
    path, wts = helpers.make_net_file(weights_dir, width, is_fc, seed=11, gain=helpers.GAIN[(width, is_fc)])
    engine.load_net(path)
    engine.set_precision('bf16x3')
    images = _image_set(1, 96, 128)
    rows, cols, idx = _blocks(images, width, 333)
    fused = engine.predict_image_blocks(width, is_fc, images, rows, cols, idx)['predictions_float32']
    _, _, _, (above, left, flat, _) = helpers.oracle_predict_blocks(wts, width, is_fc, images, idx, rows, cols)
    direct = engine.predict_batch(width, is_fc, flat if is_fc else above, None if is_fc else left)
    assert direct.shape == (len(rows), width, width, 1)
    numpy.testing.assert_array_equal(direct[..., 0], fused)


def test_empty_and_single_inputs(engine, weights_dir):
    path, wts = helpers.make_net_file(weights_dir, 8, True, seed=1)
    engine.load_net(path)
    images = _image_set(1, 64, 64)
    empty = numpy.zeros(0, dtype=numpy.int32)
    out = engine.predict_image_blocks(8, True, images, empty, empty)
    assert out['predictions_float32'].shape == (0, 8, 8)
    assert engine.predict_batch(8, True, numpy.zeros((0, 320), dtype=numpy.float32)).shape == (0, 8, 8, 1)
    one = engine.predict_image_blocks(8, True, images, numpy.array([8], dtype=numpy.int32), numpy.array([16], dtype=numpy.int32))
    pred, u8, _, _ = helpers.oracle_predict_blocks(wts, 8, True, images, numpy.zeros(1, int), [8], [16])
    helpers.check_parity(one['predictions_float32'], pred, one['predictions_uint8'], u8)


@pytest.mark.parametrize('width,is_fc', [(8, True), (16, False)])
def test_rebatching_and_repeat_are_bit_identical(engine, weights_dir, width, is_fc):
    """Fixed reduction order: a block's prediction does not depend on the batch it is computed in."""
    path, _ = helpers.make_net_file(weights_dir, width, is_fc, seed=21, gain=helpers.GAIN[(width, is_fc)])
    engine.load_net(path)
    engine.set_precision('bf16x3')
    images = _image_set(3, 96, 160)
    rows, cols, idx = _blocks(images, width, 700)
    a = engine.predict_image_blocks(width, is_fc, images, rows, cols, idx)
    b = engine.predict_image_blocks(width, is_fc, images, rows, cols, idx)
    numpy.testing.assert_array_equal(a['predictions_float32'], b['predictions_float32'])
    sel = numpy.arange(5, len(rows), 7)
    c = engine.predict_image_blocks(width, is_fc, images, rows[sel], cols[sel], idx[sel])
    numpy.testing.assert_array_equal(c['predictions_float32'], a['predictions_float32'][sel])
    numpy.testing.assert_array_equal(c['predictions_uint8'], a['predictions_uint8'][sel])
    numpy.testing.assert_array_equal(c['psnrs'], a['psnrs'][sel])


def test_unloaded_net_is_an_error(engine):
    from context_adaptive_neural_network_based_prediction_b200 import PnnError
    with pytest.raises(PnnError):
        engine.predict_batch(16, True, numpy.zeros((1, 1280), dtype=numpy.float32))


def test_shim_entry_points(weights_dir):
    """reference pnn/PredictionNeuralNetwork.py + pnn/batching.py call pattern (comparing_pnn_...py:576, 241)."""
    from context_adaptive_neural_network_based_prediction_b200.pnn.PredictionNeuralNetwork import PredictionNeuralNetwork
    from context_adaptive_neural_network_based_prediction_b200.pnn import batching
    from oracle import nets
    path, wts = helpers.make_net_file(weights_dir, 8, True, seed=31, gain=1.5)
    predictor = PredictionNeuralNetwork(10, 8, True)
    predictor.initialization(None, path)
    flat = numpy.random.default_rng(0).uniform(-118., 137., (40, 320)).astype(numpy.float32)
    out = batching.predict_by_batch_via_pnn((flat,), None, predictor, 10)
    assert out.shape == (40, 8, 8, 1) and out.dtype == numpy.float32
    helpers.check_parity(out, nets.forward_fc(wts, flat))
    with pytest.raises(ValueError):
        batching.predict_by_batch_via_pnn((flat[:33],), None, predictor, 10)
    conv = PredictionNeuralNetwork(10, 16, False)
    assert conv.strides_branch == (2, 1, 2, 1) and conv.is_fully_connected is False


@pytest.mark.gpu
def test_async_image_blocks_equal_sync(engine, weights_dir):
    """pnn_predict_image_blocks_async + pnn_synchronize: same bits as the synchronous call, several calls in flight
    (two widths, the same width twice, more calls than input sets)."""
    images = numpy.stack([helpers.synthetic_image(96, 160, s) for s in range(3)])
    jobs = []
    for width, is_fc in ((4, True), (16, False), (4, True), (8, True), (16, False)):
        path, _ = helpers.make_net_file(weights_dir, width, is_fc, seed=width)
        engine.load_net(path)
        rows, cols = helpers.grid_blocks(96, 160, width)
        rows, cols = numpy.tile(rows, 3), numpy.tile(cols, 3)
        idx = numpy.repeat(numpy.arange(3, dtype=numpy.int32), len(rows) // 3)
        if len(jobs) == 2:                                  # a different block list for the repeated width
            rows, cols, idx = rows[::2].copy(), cols[::2].copy(), idx[::2].copy()
        jobs.append((width, is_fc, rows, cols, idx))
    sync = [engine.predict_image_blocks(w, fc, images, r, c, i) for w, fc, r, c, i in jobs]
    asyn = [engine.predict_image_blocks(w, fc, images, r, c, i, wait=False) for w, fc, r, c, i in jobs]
    engine.synchronize()
    for a, b in zip(sync, asyn):
        for key in ('predictions_float32', 'predictions_uint8', 'psnrs'):
            numpy.testing.assert_array_equal(a[key], b[key])
