"""CPU tests: the oracle against the reference's own known answers and the committed golden fixtures."""
import ctypes
import os

import numpy
import pytest
import torch

from oracle import context, epilogue, nets

MEAN = 117.8952234192841

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ramp_case(width):
    # reference hevc/hm_common/c++/source_test/tests.cpp:262-316 (W in {4, 8}) and :449-487 (W = 16)
    if width == 16:
        height, stride, orow, ocol = 56, 60, 20, 18
    else:
        height, stride, orow, ocol = 32, 40, 10, 12
    plane = numpy.arange(height * stride, dtype=numpy.int32)
    units = width // 2
    return plane, stride, orow * stride + ocol, units


def _runs(values):
    """Compress a sequence into the 'a -> b' runs the reference's tests print."""
    out, start, prev = [], values[0], values[0]
    for v in values[1:]:
        if v != prev + 1:
            out.append((start, prev))
            start = v
        prev = v
    out.append((start, prev))
    return out


def _extractors():
    """The numpy restatement, the C restatement and (when built) the reference itself."""
    def via_numpy(plane, stride, origin, flags, n_avail, units, width, mean):
        return context.extract_context_portions_hm(plane, stride, origin, flags, n_avail, 4, 4, units, units, width, mean)

    out = [('numpy', via_numpy)]
    for name, path, sym, flag_t in (('c', 'oracle/_build/libpnn_oracle.so', 'oracle_extract_context_portions', numpy.uint8),
                                    ('reference', 'oracle/_ref/libextract_ref.so', 'ref_extract_context_portions', numpy.bool_)):
        full = os.path.join(ROOT, path)
        if not os.path.exists(full):
            continue
        fn = getattr(ctypes.CDLL(full), sym)
        fn.restype = ctypes.c_int

        def via_lib(plane, stride, origin, flags, n_avail, units, width, mean, fn=fn, flag_t=flag_t):
            above = numpy.zeros(3 * width * width, dtype=numpy.float32)
            left = numpy.zeros(2 * width * width, dtype=numpy.float32)
            fl = numpy.ascontiguousarray(flags, dtype=flag_t)
            code = fn(ctypes.c_void_p(plane.ctypes.data + 4 * origin), above.ctypes.data_as(ctypes.c_void_p),
                      left.ctypes.data_as(ctypes.c_void_p), fl.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n_avail),
                      4, 4, units, units, width, width, stride, ctypes.c_float(mean))
            return code, above, left
        out.append((name, via_lib))
    return out


@pytest.mark.parametrize('name,extract', _extractors())
def test_extract_kat_width_4(name, extract):
    """reference tests.cpp:340-343, 387-390, 434-437 (expected strings for W = 4)."""
    plane, stride, origin, units = _ramp_case(4)
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    code, above, left = extract(plane, stride, origin, flags, 2 * units + 1, units, 4, 0.)
    assert code == 0
    flat = numpy.concatenate([above, left]).astype(int).tolist()
    assert _runs(flat) == [(248, 259), (288, 299), (328, 339), (368, 379), (408, 411), (448, 451), (488, 491), (528, 531),
                           (568, 571), (608, 611), (648, 651), (688, 691)]
    # 2nd test: bottom-most left unit unavailable -> "... 528 -> 531 {16 times zero}"
    flags[0] = 0
    code, above, left = extract(plane, stride, origin, flags, 2 * units, units, 4, 0.)
    flat = numpy.concatenate([above, left]).astype(int).tolist()
    assert flat[-16:] == [0] * 16
    assert _runs(flat[:-16]) == [(248, 259), (288, 299), (328, 339), (368, 379), (408, 411), (448, 451), (488, 491), (528, 531)]
    # 3rd test: right-most above unit unavailable -> "248 -> 255 0 0 0 0 288 -> 295 0 0 0 0 ..."
    flags[0] = 1
    flags[2 * units] = 0
    code, above, left = extract(plane, stride, origin, flags, 2 * units, units, 4, 0.)
    a = above.astype(int).reshape(4, 12)
    for r, start in enumerate((248, 288, 328, 368)):
        assert a[r].tolist() == list(range(start, start + 8)) + [0, 0, 0, 0]
    assert _runs(left.astype(int).tolist()) == [(408 + 40 * i, 411 + 40 * i) for i in range(8)]


@pytest.mark.parametrize('name,extract', _extractors())
def test_extract_kat_width_8(name, extract):
    """reference tests.cpp:344-347, 391-394, 438-441 (expected strings for W = 8)."""
    plane, stride, origin, units = _ramp_case(8)
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    code, above, left = extract(plane, stride, origin, flags, 2 * units + 1, units, 8, 0.)
    assert code == 0
    assert _runs(above.astype(int).tolist()) == [(84 + 40 * i, 107 + 40 * i) for i in range(8)]
    assert _runs(left.astype(int).tolist()) == [(404 + 40 * i, 411 + 40 * i) for i in range(16)]
    flags[0] = 0
    code, above, left = extract(plane, stride, origin, flags, 2 * units, units, 8, 0.)
    assert left.astype(int).tolist()[-32:] == [0] * 32
    assert _runs(left.astype(int).tolist()[:-32]) == [(404 + 40 * i, 411 + 40 * i) for i in range(12)]
    flags[0] = 1
    flags[2 * units] = 0
    code, above, left = extract(plane, stride, origin, flags, 2 * units, units, 8, 0.)
    a = above.astype(int).reshape(8, 24)
    for r in range(8):
        assert a[r].tolist() == list(range(84 + 40 * r, 104 + 40 * r)) + [0] * 4


@pytest.mark.parametrize('name,extract', _extractors())
def test_extract_kat_width_16(name, extract):
    """reference tests.cpp:518, 532, 573, 587, 629 (expected strings for W = 16)."""
    plane, stride, origin, units = _ramp_case(16)
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    code, above, left = extract(plane, stride, origin, flags, 2 * units + 1, units, 16, 0.)
    assert code == 0
    assert _runs(above.astype(int).tolist()) == [(242 + 60 * i, 289 + 60 * i) for i in range(16)]
    assert _runs(left.astype(int).tolist()) == [(1202 + 60 * i, 1217 + 60 * i) for i in range(32)]
    flags[0] = flags[1] = 0
    code, above, left = extract(plane, stride, origin, flags, 2 * units - 1, units, 16, 0.)
    assert _runs(above.astype(int).tolist()) == [(242 + 60 * i, 289 + 60 * i) for i in range(16)]
    assert left.astype(int).tolist()[-128:] == [0] * 128
    assert _runs(left.astype(int).tolist()[:-128]) == [(1202 + 60 * i, 1217 + 60 * i) for i in range(24)]
    flags[0] = flags[1] = 1
    flags[2 * units] = 0
    code, above, left = extract(plane, stride, origin, flags, 2 * units, units, 16, 0.)
    a = above.astype(int).reshape(16, 48)
    for r in range(16):
        assert a[r].tolist() == list(range(242 + 60 * r, 286 + 60 * r)) + [0] * 4


@pytest.mark.parametrize('name,extract', _extractors())
def test_extract_error_codes(name, extract):
    """reference extraction_context.cpp:43-47 and :133-138 return -1."""
    plane, stride, origin, units = _ramp_case(4)
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    assert extract(plane, stride, origin, flags, 0, units, 4, 0.)[0] == -1
    flags[units] = 0
    assert extract(plane, stride, origin, flags, 2 * units, units, 4, 0.)[0] == -1


def test_extract_matches_reference_fixtures(golden_dir):
    """numpy and C restatements against outputs of the reference's own compiled function (tests/golden/extract_ref.npz)."""
    data = numpy.load(os.path.join(golden_dir, 'extract_ref.npz'))
    for name, extract in _extractors():
        for i in range(int(data['n_cases'][0])):
            p = 'c%d_' % i
            width, orow, ocol, n_avail = [int(v) for v in data[p + 'meta']]
            plane = data[p + 'plane'].astype(numpy.int32)
            stride = plane.shape[1]
            code, above, left = extract(numpy.ascontiguousarray(plane).ravel(), stride, orow * stride + ocol, data[p + 'flags'],
                                        n_avail, 2 * width // 4, width, float(data[p + 'mean'][0]))
            assert code == 0
            numpy.testing.assert_array_equal(above, data[p + 'above'], err_msg='%s case %d above' % (name, i))
            numpy.testing.assert_array_equal(left, data[p + 'left'], err_msg='%s case %d left' % (name, i))


def test_numpy_path_equals_hm_path_when_available():
    """sets/common.py slicing + preprocessing and extraction_context.cpp agree on the same pixels and masks."""
    rng = numpy.random.default_rng(3)
    for width in (4, 8, 16):
        img = rng.integers(0, 256, (4 * width, 5 * width)).astype(numpy.uint8)
        r, c = width + 1, width + 2
        for mask_w, mask_h in ((0, 0), (4, 0), (width, 4), (width, width)):
            a, l, _, va, vl = context.extract_portions(img, width, r - width, c - width)
            pa, pl = context.preprocess(a, l, 117.25, mask_w, mask_h, va, vl)
            units = 2 * width // 4
            flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
            if mask_w:
                flags[2 * units + 1 - mask_w // 4:] = 0
            if mask_h:
                flags[:mask_h // 4] = 0
            code, above, left = context.extract_context_portions_hm(img.astype(numpy.int32).ravel(), img.shape[1],
                                                                    r * img.shape[1] + c, flags, int(flags.sum()), 4, 4,
                                                                    units, units, width, 117.25)
            assert code == 0
            numpy.testing.assert_array_equal(above.reshape(width, 3 * width), pa)
            numpy.testing.assert_array_equal(left.reshape(2 * width, width), pl)


def test_preprocess_mask_validation():
    """reference sets/common.py:444-447."""
    a = numpy.zeros((8, 24), dtype=numpy.uint8)
    l = numpy.zeros((16, 8), dtype=numpy.uint8)
    for bad in ((3, 0), (0, 12), (-4, 0)):
        with pytest.raises(ValueError):
            context.preprocess(a, l, 0., bad[0], bad[1])


def test_layers_fast_vs_loops():
    rng = numpy.random.default_rng(0)
    for (k, s, cin, cout, h, w) in ((3, 1, 3, 4, 4, 12), (5, 2, 2, 3, 8, 6), (5, 2, 1, 4, 16, 48), (3, 1, 5, 2, 8, 4)):
        x = rng.standard_normal((2, h, w, cin)).astype(numpy.float32)
        wt = rng.standard_normal((k, k, cin, cout)).astype(numpy.float32)
        b = rng.standard_normal(cout).astype(numpy.float32)
        fast = nets.conv2d_same(torch.from_numpy(x), torch.from_numpy(wt), torch.from_numpy(b), s).numpy()
        numpy.testing.assert_allclose(fast, nets.conv2d_same_loops(x, wt, b, s), atol=1e-5)
        wt = rng.standard_normal((k, k, cout, cin)).astype(numpy.float32)
        fast = nets.tconv2d_same(torch.from_numpy(x), torch.from_numpy(wt), torch.from_numpy(b), s).numpy()
        numpy.testing.assert_allclose(fast, nets.tconv2d_same_loops(x, wt, b, s), atol=1e-5)


def test_layer_output_shapes():
    """reference test_pnn.py:90-114 ([2,3,2,12] for a stride-2 conv of [2,6,4,.]), :577-601 ([2,9,6,12]), :26-41 ([5,8,16,128])."""
    x = torch.zeros(2, 6, 4, 3)
    assert tuple(nets.conv2d_same(x, torch.zeros(5, 5, 3, 12), torch.zeros(12), 2).shape) == (2, 3, 2, 12)
    x = torch.zeros(2, 3, 2, 4)
    assert tuple(nets.tconv2d_same(x, torch.zeros(7, 7, 12, 4), torch.zeros(12), 3).shape) == (2, 9, 6, 12)
    # branch with strides (2, 1, 2, 1) on [5, 32, 64, 1] -> [5, 8, 16, 128]
    from context_adaptive_neural_network_based_prediction_b200 import weights as W
    wts = W.init_weights(16, False, 0)
    x = torch.zeros(5, 32, 64, 1)
    for i, s in enumerate((2, 1, 2, 1)):
        p = 'convolutional/branch_above/convolution_%d/' % i
        x = nets.conv2d_same(x, torch.from_numpy(wts[p + 'weights']), torch.from_numpy(wts[p + 'biases']), s)
    assert tuple(x.shape) == (5, 8, 16, 128)


def test_merger_single_channel_property():
    """reference test_pnn.py:43-88: only channel 0 of example 0 non-zero -> only that output map non-zero."""
    rng = numpy.random.default_rng(1)
    c = 6
    in0 = numpy.zeros((2, 4, 12, c), dtype=numpy.float32)
    in1 = numpy.zeros((2, 8, 4, c), dtype=numpy.float32)
    in0[0, :, :, 0] = 1.
    in1[0, :, :, 0] = 1.
    w = rng.standard_normal((c, 80, 16)).astype(numpy.float32)
    out = nets.channelwise_merger(torch.from_numpy(in0), torch.from_numpy(in1), torch.from_numpy(w), torch.zeros(c, 16)).numpy()
    assert numpy.abs(out[0, :, :, 0]).min() > 0.
    assert numpy.abs(out[0, :, :, 1:]).max() == 0. and numpy.abs(out[1]).max() == 0.
    numpy.testing.assert_allclose(out, nets.merger_loops(in0, in1, w, numpy.zeros((c, 16), dtype=numpy.float32)), atol=1e-5)


def test_leaky_relu_slope():
    """reference test_pnn.py:235-256 / pnn/tfutils.py:192."""
    x = torch.tensor([-10., -1., 0., 2.])
    numpy.testing.assert_allclose(nets.leaky_relu(x).numpy(), [-1., -0.1, 0., 2.], rtol=1e-6)


def test_macs_per_prediction():
    """SURVEY.md section 2 table."""
    expect = {(4, True): 2995200, (8, True): 3340800, (16, True): 4723200, (4, False): 953344, (8, False): 3774464,
              (16, False): 48750592, (32, False): 273317888, (64, False): 1180696576}
    for (w, fc), v in expect.items():
        assert nets.macs_per_prediction(w, fc) == v


def test_epilogues():
    """HM: clip then round half away (TComPrediction.cpp:632); Python: clip then round half even (tools/tools.py:49)."""
    p = numpy.array([-300., -0.5, 0.5, 1.5, 2.5, 254.5, 300., 10.4999], dtype=numpy.float32)
    assert epilogue.epilogue_hm(p, 0.).tolist() == [0, 0, 1, 2, 3, 255, 255, 10]
    assert epilogue.epilogue_numpy(p, 0.).tolist() == [0, 0, 0, 2, 2, 254, 255, 10]
    lib = ctypes.CDLL(os.path.join(ROOT, 'oracle/_build/libpnn_oracle.so'))
    out_i = numpy.zeros(p.size, dtype=numpy.int32)
    out_u = numpy.zeros(p.size, dtype=numpy.uint8)
    lib.oracle_epilogue_hm(p.ctypes.data_as(ctypes.c_void_p), p.size, ctypes.c_float(0.), out_i.ctypes.data_as(ctypes.c_void_p))
    lib.oracle_epilogue_numpy(p.ctypes.data_as(ctypes.c_void_p), p.size, ctypes.c_float(0.), out_u.ctypes.data_as(ctypes.c_void_p))
    assert out_i.tolist() == [0, 0, 1, 2, 3, 255, 255, 10] and out_u.tolist() == [0, 0, 0, 2, 2, 254, 255, 10]


def test_psnr_formula():
    """reference tools/tools.py:364-401."""
    a = numpy.full((4, 4), 10, dtype=numpy.uint8)
    b = numpy.full((4, 4), 12, dtype=numpy.uint8)
    assert abs(epilogue.psnr(a, b) - 10. * numpy.log10(255. ** 2 / (4. + 1.e-6))) < 1e-12
    assert abs(epilogue.psnr(a, a) - 10. * numpy.log10(255. ** 2 / 1.e-6)) < 1e-9


def test_fc_torch_vs_c_restatement():
    from context_adaptive_neural_network_based_prediction_b200 import weights as W
    lib = ctypes.CDLL(os.path.join(ROOT, 'oracle/_build/libpnn_oracle.so'))
    wts = W.init_weights(4, True, 5, bias_std=0.05, gain=1.5)
    x = numpy.random.default_rng(0).uniform(-118, 137, (7, 80)).astype(numpy.float32)
    ref = nets.forward_fc(wts, x).reshape(7, 16)
    cur = x
    for i in range(4):
        w = wts['fully_connected/weights_%d' % i]
        b = wts['fully_connected/biases_%d' % i]
        y = numpy.zeros((7, w.shape[1]), dtype=numpy.float32)
        lib.oracle_fc_layer(cur.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                            7, w.shape[0], w.shape[1], int(i != 3), y.ctypes.data_as(ctypes.c_void_p))
        cur = y
    numpy.testing.assert_allclose(cur, ref, atol=2e-4)


def test_golden_real_checkpoints(golden_dir):
    """The oracle reproduces the committed outputs for the two pretrained checkpoints the reference ships."""
    from context_adaptive_neural_network_based_prediction_b200 import weights as W
    from helpers import MEAN
    img = numpy.load(os.path.join(golden_dir, 'cliff_luma.npy'))
    gold = numpy.load(os.path.join(golden_dir, 'conv_real.npz'))
    for width in (4, 8):
        w_file, is_fc, wts = W.load_flat(os.path.join(golden_dir, 'conv%d_single.pnnw' % width))
        assert (w_file, is_fc) == (width, False)
        rows, cols = gold['rows_%d' % width], gold['cols_%d' % width]
        for masks in ((0, 0), (4, 4)):
            above, left, _, _ = context.gather_image_blocks(img[None], numpy.zeros(len(rows), int), rows, cols, width, MEAN, *masks)
            pred = nets.forward_conv(wts, above, left)[..., 0]
            tag = '%d_m%d%d' % (width, masks[0], masks[1])
            numpy.testing.assert_allclose(pred, gold['pred_' + tag], atol=2e-3)
            assert (epilogue.epilogue_numpy(pred, MEAN) == gold['u8_' + tag]).mean() > 0.9995


def test_offline_gather_equals_the_reference_numpy_path(golden_dir):
    """oracle.context.gather_image_blocks against outputs of the reference's OWN sets/common.py
    (extract_context_portions_targets_from_channels_plus_preprocessing, imported unmodified by
    tests/golden/make_reference_gather_golden.py): slicing, float32 mean subtraction, masks and FC flattening agree bit
    for bit, for widths 4 / 8 / 16, four mask pairs, both layouts.  The reference orders its outputs image-major."""
    import os
    from oracle import context
    g = numpy.load(os.path.join(golden_dir, 'reference_gather.npz'))
    images = g['images']
    n_img = images.shape[0]
    for width in (4, 8, 16):
        rows, cols = g['rows_%d' % width], g['cols_%d' % width]       # first pixel of the above portion = context anchor
        idx = numpy.repeat(numpy.arange(n_img), len(rows))
        r, c = numpy.tile(rows, n_img) + width, numpy.tile(cols, n_img) + width
        for masks in ((0, 0), (4, 0), (0, width), (width, 4)):
            above, left, flat, targets = context.gather_image_blocks(images, idx, r, c, width, MEAN, masks[0], masks[1])
            tag = '%d_%d_%d_' % (width, masks[0], masks[1])
            numpy.testing.assert_array_equal(flat, g['flat_' + tag + 'fc'])
            numpy.testing.assert_array_equal(above, g['above_' + tag + 'conv'])
            numpy.testing.assert_array_equal(left, g['left_' + tag + 'conv'])
            want_targets = targets.astype(numpy.float32)[..., None] - numpy.float32(MEAN)
            numpy.testing.assert_array_equal(want_targets, g['target_' + tag + 'fc'])
            numpy.testing.assert_array_equal(want_targets, g['target_' + tag + 'conv'])


@pytest.mark.parametrize('width', [4, 8])
def test_torch_oracle_against_float64_loop_goldens(golden_dir, width):
    """The torch (oneDNN) oracle against the committed float64 goldens of the literal loop implementation
    (tests/golden/make_loop_goldens.py) on the two pretrained checkpoints the reference ships, 240 real-image blocks each,
    outputs spanning +-100 pixel units: two implementations that share no layer code agree to 5e-4."""
    import os
    from context_adaptive_neural_network_based_prediction_b200 import weights as W
    g = numpy.load(os.path.join(golden_dir, 'loop_real.npz'))
    img = numpy.load(os.path.join(golden_dir, 'cliff_luma.npy'))
    _, _, wts = W.load_flat(os.path.join(golden_dir, 'conv%d_single.pnnw' % width))
    rows, cols = g['rows_%d' % width], g['cols_%d' % width]
    assert len(rows) >= 200
    for masks in ((0, 0), (4, 4)):
        above, left, _, _ = context.gather_image_blocks(img[None], numpy.zeros(len(rows), int), rows, cols, width, MEAN, masks[0], masks[1])
        pred = nets.forward_conv(wts, above, left)[..., 0]
        gold = g['pred_%d_m%d%d' % (width, masks[0], masks[1])]
        assert numpy.abs(gold).max() > 50.
        assert numpy.abs(pred - gold).max() <= 5e-4


def test_cpu_baseline_backend_matches_the_oracle(tmp_path):
    """oracle/_ref/libpnn_ref.so -- the libtorch-CPU backend behind the C ABI that stands in for the reference's TensorFlow-CPU HM
    build (baseline leg of configs[3]) -- computes the oracle's predictions: same layers, in process, batch of one."""
    import os
    import helpers
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path_lib = os.path.join(root, 'oracle', '_ref', 'libpnn_ref.so')
    if not os.path.exists(path_lib):
        pytest.skip('oracle/_ref/libpnn_ref.so has not been built (needs the libtorch headers)')
    lib = ctypes.CDLL(path_lib)
    lib.pnn_create.argtypes = [ctypes.c_char_p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.pnn_load_net.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    lib.pnn_predict_hm_context.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.pnn_destroy.argtypes = [ctypes.c_void_p]
    h = ctypes.c_void_p()
    assert lib.pnn_create(None, ctypes.c_float(MEAN), 22, 0, ctypes.byref(h)) == 0
    assert lib.pnn_create(None, ctypes.c_float(MEAN), 0, 0, ctypes.byref(ctypes.c_void_p())) == -1     # QP must be positive
    try:
        for width, is_fc in ((4, True), (8, True), (16, False)):
            path, wts = helpers.make_net_file(str(tmp_path), width, is_fc, seed=width, gain=helpers.GAIN[(width, is_fc)])
            assert lib.pnn_load_net(h, path.encode()) == 0
            rng = numpy.random.default_rng(width)
            above = rng.normal(0., 40., (width, 3 * width)).astype(numpy.float32)
            left = rng.normal(0., 40., (2 * width, width)).astype(numpy.float32)
            out = numpy.zeros((width, width), dtype=numpy.float32)
            if is_fc:
                flat = numpy.concatenate([above.ravel(), left.ravel()])
                assert lib.pnn_predict_hm_context(h, width, flat.ctypes.data, None, out.ctypes.data) == 0
                ref = nets.forward_fc(wts, flat[None])[0, :, :, 0]
            else:
                assert lib.pnn_predict_hm_context(h, width, above.ctypes.data, left.ctypes.data, out.ctypes.data) == 0
                ref = nets.forward_conv(wts, above[None, :, :, None], left[None, :, :, None])[0, :, :, 0]
            assert numpy.abs(ref).max() > 1.
            assert numpy.abs(out - ref).max() <= 1e-4
        assert lib.pnn_predict_hm_context(h, 32, above.ctypes.data, left.ctypes.data, out.ctypes.data) == -1   # no such net
    finally:
        lib.pnn_destroy(h)


def test_epilogue_and_metrics_against_the_reference_functions(golden_dir):
    """tests/golden/tools_ref.npz holds the outputs of the reference's own `cast_float_to_uint8`, `compute_psnr`
    (tools/tools.py:17-49, 364-401) and `compute_performance_neural_network_vs_hevc_best_mode`
    (comparing_pnn_ipfcns_hevc_best_mode.py:39-88), imported unmodified by tests/golden/make_tools_golden.py:
    the oracle's epilogue, PSNR and win frequency reproduce them exactly (ties at .5, clipping, identical blocks, PSNR ties)."""
    g = numpy.load(os.path.join(golden_dir, 'tools_ref.npz'))
    floats = g['floats']
    numpy.testing.assert_array_equal(epilogue.epilogue_numpy(floats.astype(numpy.float32), 0.), g['cast32'])
    numpy.testing.assert_array_equal(epilogue.epilogue_numpy(floats.astype(numpy.float64), 0.), g['cast64'])
    lib = ctypes.CDLL(os.path.join(ROOT, 'oracle/_build/libpnn_oracle.so'))
    f32 = numpy.ascontiguousarray(floats, dtype=numpy.float32)
    out_u = numpy.zeros(f32.size, dtype=numpy.uint8)
    lib.oracle_epilogue_numpy(f32.ctypes.data_as(ctypes.c_void_p), f32.size, ctypes.c_float(0.), out_u.ctypes.data_as(ctypes.c_void_p))
    numpy.testing.assert_array_equal(out_u, g['cast32'])
    psnrs = numpy.array([epilogue.psnr(a, b) for a, b in zip(g['pairs_a'], g['pairs_b'])])
    numpy.testing.assert_allclose(psnrs, g['psnrs'], rtol=0., atol=1e-12)
    psnrs_nn, frequency = epilogue.performance_vs_baseline(g['targets'][..., 0], g['preds'][..., 0], g['base'])
    numpy.testing.assert_allclose(psnrs_nn, g['psnrs_nn'], rtol=0., atol=1e-12)
    assert frequency == float(g['frequency'][1])                     # (one PSNR tie in the data: a tie is not a win)
    from context_adaptive_neural_network_based_prediction_b200 import offline
    wins = (psnrs_nn - g['base'] > 0.).astype(numpy.uint8)
    assert offline.reduce_statistics(psnrs_nn, wins)['frequency_win_pnn'] == float(g['frequency'][1])
