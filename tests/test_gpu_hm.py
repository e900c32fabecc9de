"""GPU tests of the in-loop (HM) call pair pnn_set_context / pnn_predict_hm against the oracle."""
import numpy
import pytest

import helpers
from helpers import MEAN

pytestmark = pytest.mark.gpu


def _oracle_hm(wts, width, plane, orow, ocol, flags, n_avail):
    from oracle import context, epilogue, nets
    stride = plane.shape[1]
    units = 2 * width // 4
    code, above, left = context.extract_context_portions_hm(plane.ravel(), stride, orow * stride + ocol, flags, n_avail, 4, 4,
                                                            units, units, width, MEAN)
    assert code == 0
    is_fc = width <= 8
    if is_fc:
        pred = nets.forward_fc(wts, numpy.concatenate([above, left])[None])
    else:
        pred = nets.forward_conv(wts, above.reshape(1, width, 3 * width, 1), left.reshape(1, 2 * width, width, 1))
    return pred[0, :, :, 0], epilogue.epilogue_hm(pred[0, :, :, 0], MEAN)


@pytest.mark.parametrize('width', [4, 8, 16, 32, 64])
def test_hm_call_pair(engine, weights_dir, width):
    """Net selection by width as TComPrediction.cpp:564; availability patterns as TComPattern.cpp:260-280."""
    is_fc = width <= 8
    path, wts = helpers.make_net_file(weights_dir, width, is_fc, seed=40 + width, gain=helpers.GAIN[(width, is_fc)])
    engine.load_net(path)
    engine.set_precision('bf16x3')
    plane = helpers.synthetic_image(3 * width + 8, 3 * width + 24, 5).astype(numpy.int32)
    orow, ocol = width + 3, width + 5
    units = 2 * width // 4
    patterns = []
    full = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    patterns.append(full)
    p = full.copy(); p[:units // 2] = 0; patterns.append(p)                    # below-left unavailable
    p = full.copy(); p[units + 1 + units // 2:] = 0; patterns.append(p)        # above-right unavailable
    p = full.copy(); p[:units // 2] = 0; p[units + 1 + units // 2:] = 0; patterns.append(p)
    p = full.copy(); p[:units] = 0; p[units + 1:] = 0; patterns.append(p)       # only above-left
    total_px, same_px = 0, 0
    for flags in patterns:
        n_avail = int(flags.sum())
        engine.set_context(width, plane, orow, ocol, flags, n_avail)
        got = engine.predict_hm(width, dst_stride=width + 5)
        again = engine.predict_hm(width)
        numpy.testing.assert_array_equal(got, again)                           # same bits on every call
        _, want = _oracle_hm(wts, width, plane, orow, ocol, flags, n_avail)
        assert numpy.abs(got - want).max() <= 1
        total_px += want.size
        same_px += int((got == want).sum())
        assert got.min() >= 0 and got.max() <= 255
    assert same_px >= 0.999 * total_px


def test_hm_error_behaviour(engine, weights_dir):
    """Same -1 conditions as extraction_context.cpp:17-47 and :133-138."""
    from context_adaptive_neural_network_based_prediction_b200 import PnnError
    path, _ = helpers.make_net_file(weights_dir, 8, True, seed=48)
    engine.load_net(path)
    plane = helpers.synthetic_image(40, 48, 1).astype(numpy.int32)
    flags = numpy.ones(9, dtype=numpy.uint8)
    with pytest.raises(PnnError):
        engine.set_context(8, plane, 11, 13, flags, 0)                          # iNumIntraNeighbor <= 0
    flags[4] = 0
    with pytest.raises(PnnError):
        engine.set_context(8, plane, 11, 13, flags, 8)                          # above-left unavailable
    with pytest.raises(PnnError):
        engine.predict_hm(8)                                                    # no staged context
    flags[4] = 1
    engine.set_context(8, plane, 11, 13, flags, 9)
    with pytest.raises(PnnError):
        engine.predict_hm(16)                                                   # width mismatch


@pytest.mark.parametrize('width', [4, 8, 16, 32])
def test_hm_context_call_matches_call_pair(engine, weights_dir, width):
    """pnn_predict_hm_context (what Session::Run does in HM, contexts already extracted on the host) runs the same
    batch-1 kernels as the pnn_set_context / pnn_predict_hm pair: same raw prediction, bit for bit after rounding."""
    from oracle import context, epilogue
    is_fc = width <= 8
    path, wts = helpers.make_net_file(weights_dir, width, is_fc, seed=60 + width, gain=helpers.GAIN[(width, is_fc)])
    engine.load_net(path)
    engine.set_precision('bf16x3')
    plane = helpers.synthetic_image(3 * width + 8, 3 * width + 24, 9).astype(numpy.int32)
    orow, ocol = width + 2, width + 7
    units = 2 * width // 4
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    flags[:units // 2] = 0
    n_avail = int(flags.sum())
    engine.set_context(width, plane, orow, ocol, flags, n_avail)
    pair = engine.predict_hm(width)
    stride = plane.shape[1]
    code, above, left = context.extract_context_portions_hm(plane.ravel(), stride, orow * stride + ocol, flags, n_avail, 4, 4,
                                                            units, units, width, MEAN)
    raw = engine.predict_hm_context(width, numpy.concatenate([above, left]) if is_fc else above, None if is_fc else left)
    numpy.testing.assert_array_equal(epilogue.epilogue_hm(raw, MEAN), pair)
    pred, want = _oracle_hm(wts, width, plane, orow, ocol, flags, n_avail)
    helpers.check_parity(raw, pred)


@pytest.mark.parametrize('width', [4, 8])
def test_hm_context_call_with_trained_conv_nets(engine, golden_dir, width):
    """A convolutional net loaded for width 4 / 8 serves the flattened-context call (its portions are the two halves of
    the flattened context, TComPattern.cpp:352-353); checked with the pretrained checkpoints the reference ships."""
    import os
    from context_adaptive_neural_network_based_prediction_b200 import Engine, weights as W
    from oracle import context, nets
    eng = Engine()
    try:
        path = os.path.join(golden_dir, 'conv%d_single.pnnw' % width)
        eng.load_net(path)
        _, _, wts = W.load_flat(path)
        img = numpy.load(os.path.join(golden_dir, 'cliff_luma.npy'))
        above, left, flat, _ = context.gather_image_blocks(img[None], [0, 0, 0], [width, 40, 96], [width, 64, 120], width, MEAN, 0, 0)
        for i in range(3):
            raw = eng.predict_hm_context(width, flat[i])
            ref = nets.forward_conv(wts, above[i:i + 1], left[i:i + 1])[0, :, :, 0]
            helpers.check_parity(raw, ref)
            again = eng.predict_hm_context(width, flat[i])
            numpy.testing.assert_array_equal(raw, again)
    finally:
        eng.close()


@pytest.mark.parametrize('width', [4, 8])
def test_hm_fused_fc_kernel_equals_gemv_chain(engine, weights_dir, width):
    """The fused cooperative kernel and the CUDA-graph GEMV chain use the same arithmetic and reduction order."""
    path, _ = helpers.make_net_file(weights_dir, width, True, seed=70 + width, gain=helpers.GAIN[(width, True)])
    engine.load_net(path)
    plane = helpers.synthetic_image(3 * width + 8, 3 * width + 24, 11).astype(numpy.int32)
    units = 2 * width // 4
    results = []
    try:
        for fused in (True, False, True):
            engine.set_hm_fused(fused)
            outs = []
            for shift in range(4):
                flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
                if shift & 1:
                    flags[:units // 2] = 0
                engine.set_context(width, plane, width + 1 + shift, width + 2 + shift, flags, int(flags.sum()))
                outs.append(engine.predict_hm(width).copy())
            results.append(numpy.stack(outs))
    finally:
        engine.set_hm_fused(True)
    numpy.testing.assert_array_equal(results[0], results[1])
    numpy.testing.assert_array_equal(results[0], results[2])


@pytest.mark.parametrize('width', [16, 64])
def test_hm_split_k_matches_plain_kernels(engine, weights_dir, width):
    """In-loop conv calls: split-K with a fixed-order reduction (default) against the plain batch-1 kernels: both within the
    parity bound of the oracle, each deterministic."""
    path, wts = helpers.make_net_file(weights_dir, width, False, seed=80 + width, gain=helpers.GAIN[(width, False)])
    engine.load_net(path)
    engine.set_precision('bf16x3')
    plane = helpers.synthetic_image(3 * width + 8, 3 * width + 24, 13).astype(numpy.int32)
    units = 2 * width // 4
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    n_avail = int(flags.sum())
    pred, want = _oracle_hm(wts, width, plane, width + 3, width + 5, flags, n_avail)
    outs = {}
    try:
        for mode in (True, False):
            engine.set_hm_fused(mode)
            engine.set_context(width, plane, width + 3, width + 5, flags, n_avail)
            a = engine.predict_hm(width)
            engine.set_context(width, plane, width + 3, width + 5, flags, n_avail)
            b = engine.predict_hm(width)
            numpy.testing.assert_array_equal(a, b)
            assert numpy.abs(a - want).max() <= 1 and (a == want).mean() >= 0.999
            outs[mode] = a
    finally:
        engine.set_hm_fused(True)
    assert numpy.abs(outs[True] - outs[False]).max() <= 1


def _fc_context(width, seed):
    return numpy.random.default_rng(seed).normal(0., 30., 5 * width * width).astype(numpy.float32)


def test_persistent_fc_kernel_serves_calls_without_launches(engine, weights_dir):
    """The FC nets of widths 4 and 8 live in ONE persistent kernel: after its launch an in-loop call is a doorbell write
    and a poll, whichever of the two nets it addresses; results equal the layer-per-launch graphs bit for bit and the
    oracle within the parity bound."""
    from oracle import nets
    wts = {}
    for width in (4, 8):
        path, wts[width] = helpers.make_net_file(weights_dir, width, True, seed=170 + width, gain=helpers.GAIN[(width, True)])
        engine.load_net(path)
    engine.set_hm_fused(True)
    engine.predict_hm_context(4, _fc_context(4, 0))                     # starts the kernel
    before = engine.launch_count
    outs = []
    for i in range(40):
        width = 4 if i % 3 else 8
        outs.append((width, i, engine.predict_hm_context(width, _fc_context(width, i))))
    launched = engine.launch_count - before
    try:
        engine.set_hm_fused(False)
        for width, i, raw in outs:
            numpy.testing.assert_array_equal(raw, engine.predict_hm_context(width, _fc_context(width, i)))
            helpers.check_parity(raw, nets.forward_fc(wts[width], _fc_context(width, i)[None])[0, :, :, 0])
    finally:
        engine.set_hm_fused(True)
    assert launched == 0, 'the persistent kernel was not used (%d launches for 40 calls)' % launched


def test_interleaved_fc_and_conv_calls(engine, weights_dir):
    """The codec alternates one convolutional call with tens of FC calls: the persistent kernel leaves the SMs for the
    convolutional graph and comes back behind it; every result equals the one of the plain-graph mode."""
    nets_ = ((4, True), (8, True), (16, False))
    for width, is_fc in nets_:
        path, _ = helpers.make_net_file(weights_dir, width, is_fc, seed=180 + width, gain=helpers.GAIN[(width, is_fc)])
        engine.load_net(path)
    engine.set_precision('bf16x3')
    plane = helpers.synthetic_image(80, 96, 3).astype(numpy.int32)

    def run():
        outs = []
        for i in range(60):
            width = (16, 8, 4, 4, 4, 4)[i % 6]
            units = 2 * width // 4
            flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
            if i % 4 == 1:
                flags[:units // 2] = 0
            engine.set_context(width, plane, 20 + i % 7, 24 + i % 5, flags, int(flags.sum()))
            outs.append(engine.predict_hm(width).copy())
        return outs
    fused = run()
    again = run()
    for a, b in zip(fused, again):
        numpy.testing.assert_array_equal(a, b)
    # FC results are bit-identical across the two modes (same device functions); the convolutional ones within one level
    try:
        engine.set_hm_fused(False)
        plain = run()
    finally:
        engine.set_hm_fused(True)
    for i, (a, b) in enumerate(zip(fused, plain)):
        if (16, 8, 4, 4, 4, 4)[i % 6] <= 8:
            numpy.testing.assert_array_equal(a, b)
        else:
            assert numpy.abs(a - b).max() <= 1


@pytest.mark.parametrize('width', [4, 16])
def test_hm_cache_answers_repeated_contexts_identically(engine, weights_dir, width):
    """pnn_set_hm_cache: a context seen before is answered from host memory with the very same bits; a context that
    differs in one pixel, or in one availability flag, is computed."""
    is_fc = width <= 8
    path, _ = helpers.make_net_file(weights_dir, width, is_fc, seed=190 + width, gain=helpers.GAIN[(width, is_fc)])
    engine.load_net(path)
    plane = helpers.synthetic_image(3 * width + 8, 3 * width + 24, 21).astype(numpy.int32)
    units = 2 * width // 4
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)

    def call(p, f):
        engine.set_context(width, p, width + 3, width + 5, f, int(f.sum()))
        return engine.predict_hm(width).copy()
    try:
        engine.set_hm_cache(False)
        want = call(plane, flags)
        plane2 = plane.copy(); plane2[3, width + 6] += 9               # one pixel of the above portion
        want2 = call(plane2, flags)
        flags3 = flags.copy(); flags3[-1] = 0                          # last above-right unit unavailable
        want3 = call(plane, flags3)
        engine.set_hm_cache(True)
        h0, m0 = engine.hm_cache_stats
        got = [call(plane, flags), call(plane, flags), call(plane2, flags), call(plane, flags3), call(plane2, flags)]
        h1, m1 = engine.hm_cache_stats
    finally:
        engine.set_hm_cache(False)
    assert (h1 - h0, m1 - m0) == (2, 3)
    for g, w in zip(got, (want, want, want2, want3, want2)):
        numpy.testing.assert_array_equal(g, w)
    assert (want != want2).any() and (want != want3).any()


def test_wide_fc_net_in_loop_uses_the_batched_kernels(engine, weights_dir):
    """An FC net of width 16 (an offline comparison net, 1280 inputs, 256 outputs) asked for in-loop does not fit the
    batch-1 FC kernels (5*W*W <= 320): it must go through the batched kernels with one sample, not overflow them."""
    from oracle import nets
    from context_adaptive_neural_network_based_prediction_b200 import Engine
    eng = Engine()
    try:
        path, wts = helpers.make_net_file(weights_dir, 16, True, seed=216, gain=1.6)
        eng.load_net(path)
        for seed in range(3):
            flat = _fc_context(16, seed)
            raw = eng.predict_hm_context(16, flat)
            helpers.check_parity(raw, nets.forward_fc(wts, flat[None])[0, :, :, 0])
            numpy.testing.assert_array_equal(raw, eng.predict_hm_context(16, flat))
    finally:
        eng.close()


def test_registered_nets_load_at_first_use(weights_dir):
    """pnn_register_net validates the header at once and uploads at first use; a file that lacks a tensor is refused at
    registration (as load_graph fails at start-up, integration_prediction_neural_network.cpp:29-54)."""
    import os
    from oracle import nets
    from context_adaptive_neural_network_based_prediction_b200 import Engine, PnnError, weights as W
    eng = Engine()
    try:
        path, wts = helpers.make_net_file(weights_dir, 8, True, seed=230, gain=1.6)
        eng.register_net(path)
        before = eng.launch_count
        flat = _fc_context(8, 1)
        helpers.check_parity(eng.predict_hm_context(8, flat), nets.forward_fc(wts, flat[None])[0, :, :, 0])
        assert eng.launch_count > before
        # the same tensors under the header of a width-4 net: the table no longer matches what the net needs
        bad = os.path.join(weights_dir, 'broken_fc8.pnnw')
        data = bytearray(open(path, 'rb').read())
        data[8:12] = (4).to_bytes(4, 'little')
        open(bad, 'wb').write(bytes(data))
        with pytest.raises(PnnError, match='unexpected shape'):
            eng.register_net(bad)
        truncated = os.path.join(weights_dir, 'truncated_fc8.pnnw')
        open(truncated, 'wb').write(open(path, 'rb').read()[:1 << 20])
        with pytest.raises(PnnError):
            eng.register_net(truncated)
    finally:
        eng.close()


@pytest.mark.parametrize('width', [8, 16])
def test_lazy_context_staging_gives_the_same_prediction(engine, weights_dir, width):
    """pnn_set_context_lazy: the pixels are copied by pnn_predict_hm instead of pnn_set_context (HM extracts a context for
    every transform block and candidate mode, the NN mode needs few of them); same checks, same prediction."""
    from context_adaptive_neural_network_based_prediction_b200 import PnnError
    is_fc = width <= 8
    path, _ = helpers.make_net_file(weights_dir, width, is_fc, seed=250 + width, gain=helpers.GAIN[(width, is_fc)])
    engine.load_net(path)
    plane = helpers.synthetic_image(3 * width + 8, 3 * width + 24, 31).astype(numpy.int32)
    units = 2 * width // 4
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    flags[:units // 2] = 0
    n_avail = int(flags.sum())
    engine.set_context(width, plane, width + 2, width + 4, flags, n_avail)
    eager = engine.predict_hm(width).copy()
    try:
        engine.set_context_lazy(True)
        engine.set_context(width, plane, width + 2, width + 4, flags, n_avail)
        lazy = engine.predict_hm(width).copy()
        bad = flags.copy(); bad[units] = 0                          # above-left unavailable: refused at set_context time
        with pytest.raises(PnnError):
            engine.set_context(width, plane, width + 2, width + 4, bad, n_avail - 1)
        with pytest.raises(PnnError):
            engine.predict_hm(width)                                # the failed set_context left nothing staged
    finally:
        engine.set_context_lazy(False)
    numpy.testing.assert_array_equal(eager, lazy)


def test_deferred_creation_and_warm_up(weights_dir, tmp_path):
    """pnn_create_deferred touches no GPU until a call needs it; pnn_warm_up initialises the device and uploads the registered
    nets on a thread of the library while the caller goes on; predictions equal those of an ordinary handle."""
    from context_adaptive_neural_network_based_prediction_b200 import Engine
    paths = {}
    for width in (4, 8, 16, 32, 64):
        paths[width], _ = helpers.make_net_file(weights_dir, width, width <= 8, seed=300 + width, gain=helpers.GAIN[(width, width <= 8)])
    paths_file = str(tmp_path / 'paths.txt')
    open(paths_file, 'w').write(''.join('%d,0,0,%s\n' % (w, paths[w]) for w in paths))
    plain = Engine(paths_file=paths_file, qp_selection=22)
    lazy = Engine(paths_file=paths_file, qp_selection=22, deferred=True)
    warm = Engine(paths_file=paths_file, qp_selection=22, deferred=True)
    try:
        warm.warm_up()
        for i in range(30):                                   # calls race with the uploads of the warm-up thread
            width = (4, 8, 16, 4, 8, 32, 4, 64)[i % 8]
            ctx = _fc_context(width, i)
            args = (width, ctx) if width <= 8 else (width, ctx[:3 * width * width], ctx[3 * width * width:])
            want = plain.predict_hm_context(*args)
            numpy.testing.assert_array_equal(want, lazy.predict_hm_context(*args))
            numpy.testing.assert_array_equal(want, warm.predict_hm_context(*args))
    finally:
        plain.close(); lazy.close(); warm.close()
    unused = Engine(paths_file=paths_file, qp_selection=22, deferred=True)
    unused.close()                                            # never touched the device: nothing to tear down


@pytest.mark.parametrize('width', [16, 32])
def test_fp32_in_loop_conv_path(engine, weights_dir, width):
    """fp32 precision selects the batch-1 fp32 layers for the in-loop convolutional calls (launch_gemm_skinny: 16 x 16 output
    tiles over the whole K, no split-K): within 1e-3 of the oracle, deterministic."""
    path, wts = helpers.make_net_file(weights_dir, width, False, seed=270 + width, gain=helpers.GAIN[(width, False)])
    engine.load_net(path)
    plane = helpers.synthetic_image(3 * width + 8, 3 * width + 24, 17).astype(numpy.int32)
    units = 2 * width // 4
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    flags[units + 1 + units // 2:] = 0
    n_avail = int(flags.sum())
    from oracle import context
    stride = plane.shape[1]
    _, above, left = context.extract_context_portions_hm(plane.ravel(), stride, (width + 3) * stride + width + 5, flags, n_avail, 4, 4,
                                                         units, units, width, MEAN)
    pred, want = _oracle_hm(wts, width, plane, width + 3, width + 5, flags, n_avail)
    try:
        engine.set_precision('fp32')
        raw = engine.predict_hm_context(width, above, left)
        numpy.testing.assert_array_equal(raw, engine.predict_hm_context(width, above, left))
        assert numpy.abs(raw - pred).max() <= 1e-3
        engine.set_context(width, plane, width + 3, width + 5, flags, n_avail)
        got = engine.predict_hm(width)
        assert numpy.abs(got - want).max() <= 1 and (got == want).mean() >= 0.999
    finally:
        engine.set_precision('bf16x3')


@pytest.mark.parametrize('width', [4, 8, 16, 64])
def test_posted_request_is_collected_or_remembered(engine, weights_dir, width):
    """pnn_predict_hm_begin posts the context and returns; a pnn_predict_hm that follows collects the answer (same bits as the
    synchronous call); any other call first finishes the request and leaves its answer in the memo, where the same context is
    found again (the switch codec's RD pass after the fast pass)."""
    is_fc = width <= 8
    path, _ = helpers.make_net_file(weights_dir, width, is_fc, seed=400 + width, gain=helpers.GAIN[(width, is_fc)])
    engine.load_net(path)
    units = 2 * width // 4
    flags = numpy.ones(2 * units + 1, dtype=numpy.uint8)
    n_avail = int(flags.sum())
    planes = [helpers.synthetic_image(3 * width + 8, 3 * width + 24, 50 + i).astype(numpy.int32) for i in range(3)]
    sync = []
    for plane in planes:
        engine.set_context(width, plane, width + 2, width + 4, flags, n_avail)
        sync.append(engine.predict_hm(width).copy())
    assert not numpy.array_equal(sync[0], sync[1])
    try:
        engine.set_context_lazy(True)
        engine.set_hm_cache(True)
        hits0, misses0 = engine.hm_cache_stats
        # 1. begin ... predict: collected
        engine.set_context(width, planes[0], width + 2, width + 4, flags, n_avail)
        engine.predict_hm_begin(width)
        numpy.testing.assert_array_equal(engine.predict_hm(width), sync[0])
        # 2. begin, then ANOTHER context: the posted request is finished and remembered, the new one is computed
        engine.set_context(width, planes[1], width + 2, width + 4, flags, n_avail)
        engine.predict_hm_begin(width)
        engine.set_context(width, planes[2], width + 2, width + 4, flags, n_avail)
        numpy.testing.assert_array_equal(engine.predict_hm(width), sync[2])
        hits1, misses1 = engine.hm_cache_stats
        assert (hits1 - hits0, misses1 - misses0) == (0, 3)
        # ... and the remembered one is answered from the memo, whether asked for directly or through another begin
        engine.set_context(width, planes[1], width + 2, width + 4, flags, n_avail)
        numpy.testing.assert_array_equal(engine.predict_hm(width), sync[1])
        engine.set_context(width, planes[1], width + 2, width + 4, flags, n_avail)
        engine.predict_hm_begin(width)
        numpy.testing.assert_array_equal(engine.predict_hm(width), sync[1])
        hits2, misses2 = engine.hm_cache_stats
        assert (hits2 - hits1, misses2 - misses1) == (2, 0)
        # 3. begin, then a batched call on the same handle: no dead-lock, both right
        engine.set_hm_cache(False)
        engine.set_context(width, planes[0], width + 2, width + 4, flags, n_avail)
        engine.predict_hm_begin(width)
        engine.synchronize()
        engine.set_context(width, planes[0], width + 2, width + 4, flags, n_avail)
        numpy.testing.assert_array_equal(engine.predict_hm(width), sync[0])
        # 4. begin needs a context
        from context_adaptive_neural_network_based_prediction_b200 import PnnError
        with pytest.raises(PnnError):
            engine.predict_hm_begin(width * 2 if width < 64 else 4)
    finally:
        engine.set_hm_cache(False)
        engine.set_context_lazy(False)
