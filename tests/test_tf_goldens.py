"""TensorFlow parity mode: true TensorFlow outputs, when somebody has dumped them.

`tools/dump_tf_goldens.py` (run where TensorFlow 1.x exists) writes `tests/golden/tf_<kind><W>.{npz,pnnw}`; every
pair found is checked here — the oracle on the CPU, libpnn_cuda on the GPU — to BASELINE.json's bar
(max abs error <= 1e-2 pixel units, >= 99.9 % identical rounded pixels).  No pair is committed yet (no TensorFlow in
the image): until then the tests skip and the network arithmetic stays "parity unpinned" (DESIGN.md §6).
"""
import glob
import os

import numpy
import pytest

import helpers
from context_adaptive_neural_network_based_prediction_b200 import weights as W

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
PAIRS = sorted(p[:-4] for p in glob.glob(os.path.join(GOLDEN, 'tf_*.npz')) if os.path.exists(p[:-4] + '.pnnw'))


def _load(stem):
    z = numpy.load(stem + '.npz')
    width, is_fc, wts = W.load_flat(stem + '.pnnw')
    assert (width, int(is_fc)) == (int(z['width']), int(z['is_fc']))
    return z, width, bool(is_fc), wts


@pytest.mark.skipif(not PAIRS, reason='no TensorFlow golden has been dumped (tools/dump_tf_goldens.py)')
@pytest.mark.parametrize('stem', PAIRS or ['none'])
def test_oracle_against_tensorflow(stem):
    from oracle import epilogue, nets
    z, width, is_fc, wts = _load(stem)
    n = z['above'].shape[0]
    flat = numpy.concatenate([z['above'].reshape(n, -1), z['left'].reshape(n, -1)], axis=1)
    pred = nets.forward(wts, width, is_fc, (flat,) if is_fc else (z['above'], z['left']))[..., 0]
    helpers.check_parity(pred, z['predictions'], epilogue.epilogue_numpy(pred, helpers.MEAN),
                         epilogue.epilogue_numpy(z['predictions'], helpers.MEAN))


@pytest.mark.gpu
@pytest.mark.skipif(not PAIRS, reason='no TensorFlow golden has been dumped (tools/dump_tf_goldens.py)')
@pytest.mark.parametrize('stem', PAIRS or ['none'])
def test_engine_against_tensorflow(engine, stem):
    from oracle import epilogue
    z, width, is_fc, _ = _load(stem)
    engine.load_net(stem + '.pnnw')
    n = z['above'].shape[0]
    if is_fc:
        flat = numpy.concatenate([z['above'].reshape(n, -1), z['left'].reshape(n, -1)], axis=1)
        pred = engine.predict_batch(width, True, flat)
    else:
        pred = engine.predict_batch(width, False, z['above'], z['left'])
    pred = numpy.asarray(pred).reshape(n, width, width)
    helpers.check_parity(pred, z['predictions'], epilogue.epilogue_numpy(pred, helpers.MEAN),
                         epilogue.epilogue_numpy(z['predictions'], helpers.MEAN))
