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


@pytest.mark.parametrize('width,is_fc,n_chunks', [(4, True, 3), (32, False, 4)])
def test_multi_chunk_parity_with_ragged_tail(engine, weights_dir, width, is_fc, n_chunks):
    """A small workspace budget (C ABI pnn_set_workspace_budget) cuts the call into >= 3 chunks with a ragged tail: the
    predictions match the oracle and are bit-identical to the single-chunk run (what bench.py does at full size, where
    FC-4 runs as two chunks)."""
    path, wts = helpers.make_net_file(weights_dir, width, is_fc, seed=91 + width, gain=helpers.GAIN[(width, is_fc)])
    engine.load_net(path)
    engine.set_precision('bf16x3')
    images = _image_set(2, max(96, 4 * width), max(160, 6 * width))
    rows, cols, idx = _blocks(images, width, 1000 if is_fc else 45)
    n = len(rows)
    per_sample = {True: 24 * 1024, False: 1200 * 1024}[is_fc]               # upper bounds of the bytes per sample
    try:
        engine.set_workspace_budget(20 << 30)
        whole = engine.predict_image_blocks(width, is_fc, images, rows, cols, idx)
        engine.set_workspace_budget(max(1, n // n_chunks - 1) * per_sample)
        before = engine.launch_count
        cut = engine.predict_image_blocks(width, is_fc, images, rows, cols, idx)
        launches_cut = engine.launch_count - before
        before = engine.launch_count
        engine.set_workspace_budget(20 << 30)
        engine.predict_image_blocks(width, is_fc, images, rows, cols, idx)
        launches_whole = engine.launch_count - before
    finally:
        engine.set_workspace_budget(20 << 30)
    assert launches_cut >= 2 * launches_whole, 'the small budget did not produce several chunks'
    for key in ('predictions_float32', 'predictions_uint8', 'psnrs'):
        numpy.testing.assert_array_equal(whole[key], cut[key])
    pred, u8, psnrs, _ = helpers.oracle_predict_blocks(wts, width, is_fc, images, idx, rows, cols)
    helpers.check_parity(cut['predictions_float32'], pred, cut['predictions_uint8'], u8)


@pytest.mark.parametrize('width,is_fc', [(4, True), (8, True), (16, False), (32, False)])
def test_parity_on_the_bench_shapes(engine, weights_dir, width, is_fc):
    """BASELINE.json configs[1] shapes: 320 x 480 images, EVERY grid block of two of them, the bench's nets (no gain)."""
    path, wts = helpers.make_net_file(weights_dir, width, is_fc, seed=width, bias_std=0., gain=1.)
    engine.load_net(path)
    engine.set_precision('bf16x3')
    images = _image_set(2, 320, 480)
    rows, cols, idx = _blocks(images, width)
    assert len(rows) == 2 * (320 // width - 1) * (480 // width - 1)
    out = engine.predict_image_blocks(width, is_fc, images, rows, cols, idx)
    pred, u8, psnrs, _ = helpers.oracle_predict_blocks(wts, width, is_fc, images, idx, rows, cols)
    helpers.check_parity(out['predictions_float32'], pred, out['predictions_uint8'], u8)
    same = (out['predictions_uint8'] == u8).reshape(len(rows), -1).all(axis=1)
    numpy.testing.assert_allclose(out['psnrs'][same], psnrs[same], rtol=0, atol=1e-9)


def test_conv64_parity_on_64_blocks(engine, weights_dir):
    """BASELINE.json configs[2] net at a batch that fills several M tiles of every layer."""
    path, wts = helpers.make_net_file(weights_dir, 64, False, seed=164, gain=helpers.GAIN[(64, False)])
    engine.load_net(path)
    engine.set_precision('bf16x3')
    images = _image_set(2, 4 * 64, 10 * 64)
    rows, cols, idx = _blocks(images, 64, 64)
    assert len(rows) >= 48
    out = engine.predict_image_blocks(64, False, images, rows, cols, idx)
    pred, u8, psnrs, _ = helpers.oracle_predict_blocks(wts, 64, False, images, idx, rows, cols)
    assert numpy.abs(pred).max() > 3.
    helpers.check_parity(out['predictions_float32'], pred, out['predictions_uint8'], u8)


def test_create_with_paths_file(weights_dir, tmp_path):
    """pnn_create with the reference's paths file (`width,is_pair,0,path`): single models below QP 32, pair models from
    QP 32 on when the file lists them, an error when a width is missing (TComPrediction.cpp(substitution):145-171)."""
    from context_adaptive_neural_network_based_prediction_b200 import Engine, PnnError
    from oracle import nets
    single, pair, wts = {}, {}, {}
    for width in (4, 8, 16, 32, 64):
        is_fc = width <= 8
        single[width], wts[(width, 0)] = helpers.make_net_file(weights_dir, width, is_fc, seed=300 + width, gain=helpers.GAIN[(width, is_fc)])
        pair[width], wts[(width, 1)] = helpers.make_net_file(weights_dir, width, is_fc, seed=400 + width, gain=helpers.GAIN[(width, is_fc)])
    only_single = str(tmp_path / 'single.txt')
    open(only_single, 'w').write(''.join('%d,0,0,%s\n' % (w, single[w]) for w in single))
    both = str(tmp_path / 'pair.txt')
    open(both, 'w').write(''.join('%d,0,0,%s\n%d,1,0,%s\n' % (w, single[w], w, pair[w]) for w in single) + '\n')
    missing = str(tmp_path / 'missing.txt')
    open(missing, 'w').write(''.join('%d,0,0,%s\n' % (w, single[w]) for w in (4, 8, 16, 64)))
    rng = numpy.random.default_rng(5)
    ctx = {w: rng.normal(0., 30., 5 * w * w).astype(numpy.float32) for w in (4, 16)}

    def which(engine, width):
        """0 if the engine answers with the single model of this width, 1 if with the pair model."""
        flat = ctx[width]
        raw = engine.predict_hm_context(width, flat if width <= 8 else flat[:3 * width * width],
                                        None if width <= 8 else flat[3 * width * width:])
        errs = []
        for is_pair in (0, 1):
            if width <= 8:
                ref = nets.forward_fc(wts[(width, is_pair)], flat[None])[0, :, :, 0]
            else:
                ref = nets.forward_conv(wts[(width, is_pair)], flat[:3 * width * width].reshape(1, width, 3 * width, 1),
                                        flat[3 * width * width:].reshape(1, 2 * width, width, 1))[0, :, :, 0]
            errs.append(float(numpy.abs(raw - ref).max()))
        assert min(errs) <= 1e-2 < max(errs)
        return int(numpy.argmin(errs))

    for paths_file, qp, expected in ((only_single, 22, 0), (only_single, 37, 0), (both, 31, 0), (both, 32, 1)):
        eng = Engine(paths_file=paths_file, qp_selection=qp)
        try:
            assert which(eng, 4) == expected and which(eng, 16) == expected
        finally:
            eng.close()
    with pytest.raises(PnnError, match='width 32'):
        Engine(paths_file=missing, qp_selection=22)
    with pytest.raises(PnnError):
        Engine(paths_file=str(tmp_path / 'absent.txt'), qp_selection=22)
    with pytest.raises(PnnError, match='quantization parameter'):
        Engine(paths_file=only_single, qp_selection=0)


@pytest.mark.parametrize('width', [4, 8])
def test_real_checkpoints_against_float64_loop_goldens(engine, golden_dir, width):
    """The library against goldens that do NOT come from the torch oracle: the literal float64 loop implementation
    (tests/golden/make_loop_goldens.py) on the two pretrained checkpoints, 240 real-image blocks, outputs up to +-106."""
    engine.load_net(os.path.join(golden_dir, 'conv%d_single.pnnw' % width))
    img = numpy.load(os.path.join(golden_dir, 'cliff_luma.npy'))
    g = numpy.load(os.path.join(golden_dir, 'loop_real.npz'))
    rows, cols = g['rows_%d' % width], g['cols_%d' % width]
    for precision in ('bf16x3', 'fp32'):
        engine.set_precision(precision)
        for masks in ((0, 0), (4, 4)):
            out = engine.predict_image_blocks(width, False, img, rows, cols, masks=masks)
            gold = g['pred_%d_m%d%d' % (width, masks[0], masks[1])]
            err = float(numpy.abs(out['predictions_float32'] - gold).max())
            assert err <= TOL[precision], 'max abs error %.3e' % err
            want_u8 = numpy.clip(numpy.round(numpy.clip(gold.astype(numpy.float32) + numpy.float32(MEAN), 0., 255.)), 0, 255).astype(numpy.uint8)
            assert (out['predictions_uint8'] == want_u8).mean() >= 0.999
    engine.set_precision('bf16x3')


def test_device_pointer_calls_contain_invalid_blocks(engine, weights_dir):
    """The device-pointer entry points cannot validate their block lists: a block that leaves the image or names a missing
    image must not read out of bounds -- its PSNR is NaN (HEVC baseline: index 255), the valid blocks are unaffected."""
    import torch
    path, _ = helpers.make_net_file(weights_dir, 8, True, seed=333, gain=1.6)
    engine.load_net(path)
    engine.set_precision('bf16x3')
    images = _image_set(2, 64, 96)
    dev = torch.device('cuda', 0)
    rows = numpy.array([8, 16, 60, 8, -5, 24], dtype=numpy.int32)        # 60 + 8 > 64; -5 < 0
    cols = numpy.array([8, 40, 8, 8, 8, 90], dtype=numpy.int32)          # 90 + 8 > 96
    idx = numpy.array([0, 1, 0, 7, 1, 1], dtype=numpy.int32)             # image 7 does not exist
    valid = numpy.array([True, True, False, False, False, False])
    d_img = torch.from_numpy(images).to(dev)
    d_r, d_c, d_i = (torch.from_numpy(a).to(dev) for a in (rows, cols, idx))
    d_u8 = torch.zeros((6, 64), dtype=torch.uint8, device=dev)
    d_psnr = torch.zeros(6, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    engine.predict_image_blocks_device(8, True, d_img.data_ptr(), 2, 64, 96, d_i.data_ptr(), d_r.data_ptr(), d_c.data_ptr(), 6, (0, 0),
                                       None, d_u8.data_ptr(), d_psnr.data_ptr(), stream)
    d_best = torch.zeros(6, dtype=torch.uint8, device=dev)
    d_hpsnr = torch.zeros(6, dtype=torch.float64, device=dev)
    engine.hevc_best_mode_device(8, d_img.data_ptr(), 2, 64, 96, d_i.data_ptr(), d_r.data_ptr(), d_c.data_ptr(), 6, (0, 0),
                                 d_best.data_ptr(), d_hpsnr.data_ptr(), None, stream)
    torch.cuda.synchronize()
    psnr, hpsnr, best = d_psnr.cpu().numpy(), d_hpsnr.cpu().numpy(), d_best.cpu().numpy()
    assert numpy.isnan(psnr[~valid]).all() and numpy.isfinite(psnr[valid]).all()
    assert numpy.isnan(hpsnr[~valid]).all() and (best[~valid] == 255).all() and (best[valid] < 35).all()
    ref = engine.predict_image_blocks(8, True, images, rows[valid], cols[valid], idx[valid])
    numpy.testing.assert_array_equal(d_u8.cpu().numpy()[valid].reshape(-1, 8, 8), ref['predictions_uint8'])
    numpy.testing.assert_array_equal(psnr[valid], ref['psnrs'])


@pytest.mark.gpu
def test_driver_loop_over_training_and_test_maskings(engine, weights_dir, tmp_path):
    """offline.predict_masks: the reference's loop over PNN models (one per training masking; missing ones are skipped, the
    longest training is taken) and test maskings, PNN against the best HEVC mode; one cell checked against the oracle."""
    import pickle
    import shutil
    from context_adaptive_neural_network_based_prediction_b200 import offline
    from oracle import context, epilogue, nets
    width = 8
    images = numpy.stack([helpers.synthetic_image(96, 128, s) for s in range(2)])
    idx, rows, cols = offline.blocks_of_images(2, 96, 128, width)
    root = tmp_path / 'models'
    kept = {}
    for tag, seed, iterations in (('masks_tr_0_0', 11, (10, 500)), ('masks_tr_random', 12, (40,))):
        (root / tag).mkdir(parents=True)
        for it in iterations:
            path, wts = helpers.make_net_file(weights_dir, width, True, seed=seed + it, gain=helpers.GAIN[(width, True)])
            shutil.copyfile(path, str(root / tag / ('model_%d.pnnw' % it)))
            kept[tag] = wts                                                        # the last one = the longest training
    (root / 'masks_tr_8_8').mkdir()                                                # a directory without a model: skipped
    vis = tmp_path / 'vis'
    out = offline.predict_masks(engine, images, width, True, rows, cols, str(root), image_index=idx,
                                path_to_directory_coeffs_vis=str(vis))
    assert sorted(out) == ['masks_tr_0_0', 'masks_tr_random']
    assert sorted(out['masks_tr_random']) == ['masks_val_0_0', 'masks_val_0_8', 'masks_val_8_0', 'masks_val_8_8']
    cell = out['masks_tr_0_0']['masks_val_8_0']
    above, left, flat, targets = context.gather_image_blocks(images, idx, rows, cols, width, MEAN, 8, 0)
    ref = nets.forward(kept['masks_tr_0_0'], width, True, (flat,))[..., 0]
    ref_u8 = epilogue.epilogue_numpy(ref, MEAN)
    assert (cell['predictions_pnn_uint8'] == ref_u8).mean() >= 0.999
    psnr_ref = numpy.array([epilogue.psnr(ref_u8[i], targets[i]) for i in range(len(rows))])
    close = numpy.isclose(cell['psnrs_pnn'], psnr_ref, rtol=0., atol=1e-9)
    assert close.mean() >= 0.99                                                    # (blocks with a differently rounded pixel differ)
    assert cell['frequency_win_pnn'] == numpy.count_nonzero(cell['psnrs_pnn'] - cell['psnrs_hevc_best_mode'] > 0.) / len(rows)
    # the masking changes the prediction, and the two models differ
    assert not numpy.array_equal(cell['predictions_pnn_uint8'], out['masks_tr_0_0']['masks_val_0_0']['predictions_pnn_uint8'])
    assert not numpy.array_equal(cell['predictions_pnn_uint8'], out['masks_tr_random']['masks_val_8_0']['predictions_pnn_uint8'])
    saved = pickle.load(open(str(vis / 'masks_tr_0_0' / 'masks_val_8_0' / 'dictionary_performance.pkl'), 'rb'))
    assert saved['mean_psnr_pnn'] == cell['mean_psnr_pnn'] and saved['frequency_win_pnn'] == cell['frequency_win_pnn']


@pytest.mark.gpu
def test_workspace_shrinks_when_device_memory_is_short(engine, weights_dir):
    """Another tenant holds almost all of the GPU's memory: the library gives back the workspaces of its other nets, halves its
    chunk until the allocation fits, and returns the SAME bits as with the whole device to itself."""
    import torch
    width = 32
    path, _ = helpers.make_net_file(weights_dir, width, False, seed=77, gain=helpers.GAIN[(width, False)])
    engine.load_net(path)
    path8, _ = helpers.make_net_file(weights_dir, 8, True, seed=78, gain=helpers.GAIN[(8, True)])
    engine.load_net(path8)
    images = numpy.stack([helpers.synthetic_image(320, 480, s) for s in range(8)])
    idx, rows, cols = __import__('context_adaptive_neural_network_based_prediction_b200.offline', fromlist=['offline']).blocks_of_images(8, 320, 480, width)
    assert len(rows) == 8 * 126
    idx8, rows8, cols8 = __import__('context_adaptive_neural_network_based_prediction_b200.offline', fromlist=['offline']).blocks_of_images(8, 320, 480, 8)
    engine.predict_image_blocks(8, True, images, rows8, cols8, idx8)                  # another net of the handle owns a workspace
    engine.set_workspace_budget(1 << 40)                                              # forget any limit, start from nothing
    free, total = torch.cuda.mem_get_info()
    hog = torch.empty(max(free - (512 << 20), 1 << 20), dtype=torch.uint8, device='cuda')  # leave 0.5 GB: 1008 CONV-32 samples need > 1 GB
    launches = engine.launch_count
    try:
        tight = engine.predict_image_blocks(width, False, images, rows, cols, idx)
    finally:
        del hog
        torch.cuda.empty_cache()
    launches_tight = engine.launch_count - launches
    engine.set_workspace_budget(20 << 30)
    roomy = engine.predict_image_blocks(width, False, images, rows, cols, idx)
    assert launches_tight > engine.launch_count - launches - launches_tight          # more chunks when memory was short
    numpy.testing.assert_array_equal(tight['predictions_uint8'], roomy['predictions_uint8'])
    numpy.testing.assert_array_equal(tight['predictions_float32'], roomy['predictions_float32'])
    numpy.testing.assert_array_equal(tight['psnrs'], roomy['psnrs'])
    again = engine.predict_image_blocks(8, True, images, rows8, cols8, idx8)         # the net whose workspace was given back still works
    assert again['predictions_uint8'].shape == (len(rows8), 8, 8)
