"""CPU tests of the boundary: the C-ABI library loads and exports what include/pnn_cuda.h declares; host logic."""
import ctypes
import os
import re

import numpy
import pytest

from context_adaptive_neural_network_based_prediction_b200 import _lib, weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'pnn_cuda.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pnn_[a-z_0-9]+)\s*\(', text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert b'sm_100a' in lib.pnn_version()


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: on a box without an sm_100 GPU pnn_create returns -1 with a message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.pnn_create(None, ctypes.c_float(117.9), 22, 0, ctypes.byref(h)) == -1
    assert not h.value
    assert b'CUDA' in lib.pnn_last_error(None)
    from context_adaptive_neural_network_based_prediction_b200 import Engine, PnnError
    with pytest.raises(PnnError):
        Engine()


def test_create_rejects_non_positive_qp():
    """reference TComPrediction.cpp(substitution):129-133."""
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.pnn_create(None, ctypes.c_float(117.9), 0, 0, ctypes.byref(h)) == -1
    assert b'quantization parameter' in lib.pnn_last_error(None)


def test_null_handle_is_an_error_not_a_crash():
    lib = _lib.load()
    assert lib.pnn_load_net(None, b'x') == -1
    assert lib.pnn_predict_hm(None, 8, None, 8) == -1
    assert lib.pnn_launch_count(None) == 0


def test_param_counts_and_flat_roundtrip(tmp_path):
    """parameter counts pinned by the reference's checkpoint index files."""
    expect = {(4, True): 2998816, (8, True): 3344464, (16, True): 4727056, (4, False): 70145, (8, False): 198657,
              (16, False): 1339073, (32, False): 5622657, (64, False): 20652545}
    for (w, fc), count in expect.items():
        shapes = W.tensor_shapes(w, fc)
        assert sum(int(numpy.prod(s)) for s in shapes.values()) == count
    # FC-8 ends with a 256-byte tensor (exactly one alignment unit): its last byte must survive the padding
    for width, is_fc in ((8, False), (8, True), (4, True)):
        wts = W.init_weights(width, is_fc, 3, bias_std=0.1)
        path = str(tmp_path / ('n_%d_%d.pnnw' % (width, is_fc)))
        W.save_flat(path, width, is_fc, wts)
        assert os.path.getsize(path) % 256 == 0
        w, fc, back = W.load_flat(path)
        assert (w, fc) == (width, is_fc) and list(back) == list(wts)
        for k in wts:
            numpy.testing.assert_array_equal(back[k], wts[k])


def test_initialisers_follow_reference():
    """reference pnn/components.py:128-166 and pnn/tfutils.py:54,113-117,434-438."""
    fc = W.init_weights(8, True, 0)
    assert abs(fc['fully_connected/weights_0'].std() - 0.01) < 5e-4
    assert abs(fc['fully_connected/weights_1'].std() - 0.029) < 5e-4
    assert abs(fc['fully_connected/weights_3'].std() - 0.01) < 5e-4
    assert not fc['fully_connected/biases_2'].any()
    cv = W.init_weights(16, False, 0)
    assert abs(cv['convolutional/branch_left/convolution_0/weights'].std() - 0.01) < 2e-3
    w1 = cv['convolutional/branch_above/convolution_1/weights']
    assert abs(w1.std() - 1. / numpy.sqrt(64 * 9)) < 2e-3
    assert abs(cv['convolutional/merger/channelwise_fully_connected_merger/weights'].std() - 1. / numpy.sqrt(80)) < 2e-3
    assert abs(cv['convolutional/merger/transpose_convolution_3/weights'].std() - 0.01) < 2e-3


def test_golden_flat_binaries_match_reference_shapes(golden_dir):
    for width in (4, 8):
        w, fc, wts = W.load_flat(os.path.join(golden_dir, 'conv%d_single.pnnw' % width))
        assert {k: v.shape for k, v in wts.items()} == W.tensor_shapes(width, False)


def test_batching_shim_keeps_reference_errors():
    """reference pnn/batching.py:56-69 via tools/tools.py:403-434."""
    from context_adaptive_neural_network_based_prediction_b200.pnn import batching

    class Fake(object):
        is_fully_connected = True
        engine = None
    with pytest.raises(ValueError):
        batching.predict_by_batch_via_pnn((numpy.zeros((7, 320), dtype=numpy.float32),), None, Fake(), 10)
    with pytest.raises(ValueError):
        batching.predict_by_batch_via_pnn((numpy.zeros((10, 321), dtype=numpy.float32),), None, Fake(), 10)
    with pytest.raises(TypeError):
        batching.predict_by_batch_via_pnn((numpy.zeros((10, 320), dtype=numpy.float32),), None, Fake(), 10.)
