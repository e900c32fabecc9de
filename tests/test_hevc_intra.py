"""HEVC best-mode baseline: oracle against the reference's own compiled code and committed
outputs of it (CPU), CUDA kernel against the oracle, bit for bit (GPU)."""
import ctypes
import os

import numpy
import pytest

import helpers
from oracle import epilogue, hevc_intra

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_matches_committed_reference_outputs(golden_dir):
    data = numpy.load(os.path.join(golden_dir, 'hevc_ref.npz'))
    for i in range(int(data['n_cases'][0])):
        width = int(data['h%d_width' % i][0])
        preds = hevc_intra.predict_all_modes(data['h%d_row' % i].astype(numpy.int64), data['h%d_col' % i].astype(numpy.int64), width)
        numpy.testing.assert_array_equal(preds, data['h%d_preds' % i], err_msg='case %d (W = %d)' % (i, width))


def test_oracle_matches_compiled_reference():
    """Only where oracle/_ref was built (container with /root/reference)."""
    path = os.path.join(ROOT, 'oracle', '_ref', 'libhevc_intra_ref.so')
    if not os.path.exists(path):
        pytest.skip('oracle/_ref/libhevc_intra_ref.so not built')
    lib = ctypes.CDLL(path)
    rng = numpy.random.default_rng(5)
    for width in (4, 8, 16, 32):
        for _ in range(3):
            mask_w, mask_h = (4 * int(rng.integers(0, width // 4 + 1)) for _ in range(2))
            hp, wp = 2 * width + 1 - mask_h, 2 * width + 1 - mask_w
            pattern = rng.integers(0, 256, (hp, wp)).astype(numpy.uint8)
            mine = hevc_intra.predict_all_modes(pattern[0].astype(numpy.int64), pattern[:, 0].astype(numpy.int64), width)
            for mode in range(35):
                out = numpy.zeros(width * width, dtype=numpy.uint8)
                assert lib.ref_hevc_intraprediction(hp, wp, width, pattern.ctypes.data_as(ctypes.c_void_p),
                                                    out.ctypes.data_as(ctypes.c_void_p), mode) == 0
                numpy.testing.assert_array_equal(mine[mode].ravel(), out)
    out = numpy.zeros(16, dtype=numpy.uint8)
    pattern = numpy.zeros((9, 9), dtype=numpy.uint8)
    assert lib.ref_hevc_intraprediction(9, 9, 4, pattern.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), 35) == -1


def test_oracle_matches_the_reference_python_path(golden_dir):
    """tests/golden/hevc_python_ref.npz: the reference's own Python functions -- `extract_intra_patterns`,
    `predict_series_via_hevc_best_mode` (hevc/intraprediction/intraprediction.py:10-292), imported unmodified over the
    reference's C++ compiled unmodified (tests/golden/make_hevc_python_golden.py) -- on 16 (width, masks) cases of 12
    blocks, one of them in a flat area where modes tie: pattern, best-mode index, PSNR and prediction of the oracle agree."""
    data = numpy.load(os.path.join(golden_dir, 'hevc_python_ref.npz'))
    image = data['image']
    for i in range(int(data['n_cases'][0])):
        width = int(data['c%d_width' % i][0])
        mask_w, mask_h = (int(v) for v in data['c%d_masks' % i])
        row_refs, col_refs = data['c%d_row_refs' % i], data['c%d_col_refs' % i]
        patterns = data['c%d_patterns' % i]
        for j in range(len(row_refs)):
            first_row, first_col = hevc_intra.extract_intra_pattern(image, width, int(row_refs[j]), int(col_refs[j]), mask_w, mask_h)
            numpy.testing.assert_array_equal(first_row, patterns[j, 0, :, 0])
            numpy.testing.assert_array_equal(first_col, patterns[j, :, 0, 0])
        idx, psnrs, preds = hevc_intra.best_modes_of_blocks(image[None], numpy.zeros(len(row_refs), dtype=int), row_refs + 1, col_refs + 1,
                                                            width, mask_w, mask_h)
        numpy.testing.assert_array_equal(idx, data['c%d_indices' % i], err_msg='case %d' % i)
        numpy.testing.assert_allclose(psnrs, data['c%d_psnrs' % i], rtol=0., atol=1e-9)
        numpy.testing.assert_array_equal(preds, data['c%d_preds' % i][..., 0])


def test_best_mode_rule():
    """reference intraprediction.py:262-292: the first mode with the strictly highest PSNR wins; flat content -> planar (0)."""
    flat_row, flat_col = numpy.full(9, 77, dtype=numpy.int64), numpy.full(9, 77, dtype=numpy.int64)
    idx, psnr, pred = hevc_intra.best_mode(flat_row, flat_col, numpy.full((4, 4), 77, dtype=numpy.uint8))
    assert idx == 0 and (pred == 77).all() and abs(psnr - epilogue.psnr(pred, pred)) < 1e-9
    # a vertical stripe pattern is predicted exactly by the vertical mode (26), and by no earlier mode
    row = numpy.array([50, 10, 200, 30, 90, 60, 60, 60, 60], dtype=numpy.int64)
    col = numpy.full(9, 50, dtype=numpy.int64)
    target = numpy.tile(row[1:5].astype(numpy.uint8), (4, 1))
    target[:, 0] = numpy.clip(row[1] + ((col[1:5] - col[0]) >> 1), 0, 255)      # edge filter of mode 26
    idx, _, pred = hevc_intra.best_mode(row, col, target)
    assert idx == 26 and (pred == target).all()


@pytest.mark.gpu
@pytest.mark.parametrize('width', [4, 8, 16, 32, 64])
def test_gpu_best_mode_bit_exact(engine, width):
    images = numpy.stack([helpers.synthetic_image(max(96, 3 * width), max(128, 4 * width), s) for s in range(2)])
    rng = numpy.random.default_rng(width)
    images[1] = rng.integers(0, 256, images[1].shape).astype(numpy.uint8)          # hard content: many near-ties
    rows, cols = helpers.grid_blocks(images.shape[1], images.shape[2], width)
    # also blocks off the grid and touching the top / left image border (anchor row / column 0)
    rows = numpy.concatenate([rows, [1, 1, 7]]).astype(numpy.int32)
    cols = numpy.concatenate([cols, [1, 9, 1]]).astype(numpy.int32)
    keep = (rows + width <= images.shape[1]) & (cols + width <= images.shape[2])
    rows, cols = rows[keep][:260], cols[keep][:260]
    idx = (numpy.arange(len(rows)) % 2).astype(numpy.int32)
    for masks in ((0, 0), (4, 0), (0, width), (width, width)):
        out = engine.hevc_best_mode(width, images, rows, cols, idx, masks=masks)
        want_idx, want_psnr, want_pred = hevc_intra.best_modes_of_blocks(images, idx, rows, cols, width, masks[0], masks[1])
        numpy.testing.assert_array_equal(out['indices_hevc_best_mode'], want_idx)
        numpy.testing.assert_array_equal(out['predictions_hevc_best_mode_uint8'], want_pred)
        numpy.testing.assert_allclose(out['psnrs_hevc_best_mode'], want_psnr, rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_gpu_best_mode_equals_the_reference_python_path(engine, golden_dir):
    """The CUDA kernel against tests/golden/hevc_python_ref.npz directly (the reference's Python best-mode loop over its own
    C++, see test_oracle_matches_the_reference_python_path): indices, predictions and PSNRs, no oracle in between."""
    data = numpy.load(os.path.join(golden_dir, 'hevc_python_ref.npz'))
    image = numpy.ascontiguousarray(data['image'])[None]
    for i in range(int(data['n_cases'][0])):
        width = int(data['c%d_width' % i][0])
        masks = tuple(int(v) for v in data['c%d_masks' % i])
        rows = (data['c%d_row_refs' % i] + 1).astype(numpy.int32)
        cols = (data['c%d_col_refs' % i] + 1).astype(numpy.int32)
        out = engine.hevc_best_mode(width, image, rows, cols, numpy.zeros(len(rows), dtype=numpy.int32), masks=masks)
        numpy.testing.assert_array_equal(out['indices_hevc_best_mode'], data['c%d_indices' % i], err_msg='case %d' % i)
        numpy.testing.assert_array_equal(out['predictions_hevc_best_mode_uint8'], data['c%d_preds' % i][..., 0])
        numpy.testing.assert_allclose(out['psnrs_hevc_best_mode'], data['c%d_psnrs' % i], rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_gpu_best_mode_errors(engine):
    from context_adaptive_neural_network_based_prediction_b200 import PnnError
    img = helpers.synthetic_image(64, 64, 0)
    one = numpy.array([8], dtype=numpy.int32)
    with pytest.raises(PnnError):
        engine.hevc_best_mode(8, img, one, one, masks=(3, 0))                       # intraprediction.py:66-72
    with pytest.raises(PnnError):
        engine.hevc_best_mode(8, img, numpy.array([0], dtype=numpy.int32), one)     # no pixel above the block
    with pytest.raises(PnnError):
        engine.hevc_best_mode(12, img, one, one)
    assert engine.hevc_best_mode(8, img, one[:0], one[:0])['indices_hevc_best_mode'].shape == (0,)


@pytest.mark.gpu
def test_offline_evaluation_keys_and_win_rate(engine, golden_dir):
    """offline.evaluate_blocks = the reference's predict_mask body; checked with a trained checkpoint on a real image."""
    from context_adaptive_neural_network_based_prediction_b200 import offline, weights as W
    width = 8
    path = os.path.join(golden_dir, 'conv8_single.pnnw')
    engine.load_net(path)
    _, _, wts = W.load_flat(path)
    img = numpy.load(os.path.join(golden_dir, 'cliff_luma.npy'))
    rows, cols = helpers.grid_blocks(img.shape[0], img.shape[1], width)
    idx = numpy.zeros(len(rows), dtype=numpy.int64)
    got = offline.evaluate_blocks(engine, img[None], width, False, rows, cols)
    _, pred_u8, psnrs_pnn, _ = helpers.oracle_predict_blocks(wts, width, False, img[None], idx, rows, cols)
    _, psnrs_hevc, _ = hevc_intra.best_modes_of_blocks(img[None], idx, rows, cols, width, 0, 0)
    freq = float(numpy.count_nonzero(psnrs_pnn - psnrs_hevc > 0.)) / len(rows)
    numpy.testing.assert_allclose(got['psnrs_hevc_best_mode'], psnrs_hevc, rtol=0, atol=1e-9)
    same_px = got['predictions_pnn_uint8'].reshape(len(rows), -1) == pred_u8.reshape(len(rows), -1)
    assert same_px.mean() >= 0.999                       # BASELINE.json: >= 99.9 % identical rounded pixels
    same = same_px.all(axis=1)                           # blocks without a single rounding flip
    assert same.mean() > 0.95
    numpy.testing.assert_allclose(got['psnrs_pnn'][same], psnrs_pnn[same], rtol=0, atol=1e-9)
    assert abs(got['frequency_win_pnn'] - freq) <= (1. - same.mean()) + 1e-12
    assert set(('psnrs_pnn', 'indices_hevc_best_mode', 'psnrs_hevc_best_mode', 'mean_psnr_pnn', 'frequency_win_pnn')) <= set(got)
