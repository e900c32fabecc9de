"""Weight import paths that need no TensorFlow: the frozen-graph (binary GraphDef) reader and the exporter CLI.

The GraphDef of the test is produced by a small protobuf ENCODER written here (independent of the reader under
test), with the node kinds a frozen PNN graph holds: Placeholder, Const (tensor_content, packed float_val, single
splat value, int32 constants), ops with string / list / shape attributes (reference freezing_graph_pnn.py:129-139).
"""
import os
import struct

import numpy
import pytest

from context_adaptive_neural_network_based_prediction_b200 import weights as W


def _vi(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(field, payload):
    return _vi(field << 3 | 2) + _vi(len(payload)) + payload


def _v(field, value):
    return _vi(field << 3) + _vi(value)


def _shape(dims):
    return b''.join(_ld(2, _v(1, d)) for d in dims)


def _tensor(arr, mode):
    arr = numpy.asarray(arr)
    dtype = {numpy.dtype('float32'): 1, numpy.dtype('int32'): 3}[arr.dtype]
    msg = _v(1, dtype) + _ld(2, _shape(arr.shape))
    if mode == 'content':
        msg += _ld(4, arr.astype(arr.dtype.newbyteorder('<')).tobytes())
    elif mode == 'packed':
        msg += _ld(5, arr.astype('<f4').tobytes())
    elif mode == 'repeated':
        msg += b''.join(_vi(5 << 3 | 5) + struct.pack('<f', v) for v in arr.ravel())
    elif mode == 'splat':
        msg += _vi(5 << 3 | 5) + struct.pack('<f', float(arr.ravel()[0]))
    elif mode == 'int':
        msg += _ld(7, b''.join(_vi(int(v)) for v in arr.ravel()))
    return msg


def _attr(key, attr_value):
    return _ld(5, _ld(1, key.encode()) + _ld(2, attr_value))


def _node(name, op, inputs=(), attrs=()):
    msg = _ld(1, name.encode()) + _ld(2, op.encode())
    for i in inputs:
        msg += _ld(3, i.encode())
    for a in attrs:
        msg += a
    return _ld(1, msg)


def _const(name, arr, mode):
    arr = numpy.asarray(arr)
    dtype = 1 if arr.dtype == numpy.float32 else 3
    return _node(name, 'Const', attrs=[_attr('dtype', _v(6, dtype)), _attr('value', _ld(8, _tensor(arr, mode)))])


def make_graph(path, weights, width, is_fc):
    modes = ['content', 'packed', 'repeated']
    g = b''
    g += _node('node_flattened_context' if is_fc else 'node_portion_above', 'Placeholder',
               attrs=[_attr('dtype', _v(6, 1)), _attr('shape', _ld(7, _shape((1, 5 * width ** 2))))])
    for i, (n, a) in enumerate(weights.items()):
        g += _const(n, a, modes[i % 3] if a.size < 5000 else 'content')
        g += _node(n + '/read', 'Identity', inputs=[n], attrs=[_attr('T', _v(6, 1)), _attr('_class', _ld(1, _ld(2, b'loc:@' + n.encode())))])
    g += _const('some/int_constant', numpy.array([1, 2, 2, 1], dtype=numpy.int32), 'int')
    g += _node('conv', 'Conv2D', inputs=['a', 'b'],
               attrs=[_attr('strides', _ld(1, _ld(3, b''.join(_vi(v) for v in (1, 2, 2, 1))))), _attr('padding', _ld(2, b'SAME')),
                      _attr('use_cudnn_on_gpu', _v(5, 1))])
    g += _ld(4, _v(1, 24))                                # GraphDef.versions
    with open(path, 'wb') as f:
        f.write(g)


@pytest.mark.parametrize('width,is_fc', [(4, True), (8, False)])
def test_frozen_graph_roundtrip(tmp_path, width, is_fc):
    wts = W.init_weights(width, is_fc, seed=3, bias_std=0.1)
    graph = str(tmp_path / 'graph_output.pbtxt')
    make_graph(graph, wts, width, is_fc)
    consts, nodes = W.read_frozen_graph(graph)
    assert ('conv', 'Conv2D') in nodes and 'some/int_constant' not in consts
    out = str(tmp_path / 'net.pnnw')
    W.export_frozen_graph(graph, width, is_fc, out)
    w2, fc2, back = W.load_flat(out)
    assert (w2, bool(fc2)) == (width, is_fc)
    for n, a in wts.items():
        numpy.testing.assert_array_equal(back[n], a)


def test_frozen_graph_splat_and_errors(tmp_path):
    path = str(tmp_path / 'g.pb')
    with open(path, 'wb') as f:
        f.write(_const('fully_connected/biases_0', numpy.full((1200,), 0.25, dtype=numpy.float32), 'splat'))
    consts, _ = W.read_frozen_graph(path)
    numpy.testing.assert_array_equal(consts['fully_connected/biases_0'], numpy.full((1200,), 0.25, dtype=numpy.float32))
    with pytest.raises(ValueError, match='holds no constant named'):
        W.export_frozen_graph(path, 4, True, str(tmp_path / 'x.pnnw'))
    text = str(tmp_path / 't.pbtxt')
    with open(text, 'w') as f:
        f.write('node {\n  name: "x"\n}\n')
    with pytest.raises(ValueError, match='text-format'):
        W.read_frozen_graph(text)


def test_exporter_cli_on_reference_shaped_checkpoint(tmp_path, golden_dir):
    """`python -m <pkg>.weights --frozen-graph` end to end; result equals the committed export of the real checkpoint."""
    _, _, real = W.load_flat(os.path.join(golden_dir, 'conv4_single.pnnw'))
    graph = str(tmp_path / 'graph_output.pbtxt')
    make_graph(graph, real, 4, False)
    out = str(tmp_path / 'cli.pnnw')
    W.main(['--frozen-graph', graph, '--width', '4', '--conv', '--out', out])
    with open(out, 'rb') as a, open(os.path.join(golden_dir, 'conv4_single.pnnw'), 'rb') as b:
        assert a.read() == b.read()


def _checksum(wts):
    total = 0.
    for n in sorted(wts):                                    # std::map order = byte-wise name order
        a = wts[n].astype(numpy.float64).ravel()
        total += float(numpy.sum((numpy.arange(a.size) % 7 + 1) * a))
    return total


@pytest.mark.parametrize('width,is_fc', [(4, True), (8, True), (4, False), (8, False), (16, False), (32, False), (64, False)])
def test_library_reads_frozen_graphs_directly(tmp_path, width, is_fc):
    """The C++ GraphDef reader of libpnn_cuda (what HM's unchanged paths file exercises) against the Python reader:
    same tensors, and width / kind inferred from the constants alone.  Host-only call, no GPU."""
    from context_adaptive_neural_network_based_prediction_b200 import engine as E
    wts = W.init_weights(width, is_fc, seed=width, bias_std=0.1)
    graph = str(tmp_path / 'graph_output.pbtxt')
    make_graph(graph, wts, width, is_fc)
    flat = str(tmp_path / 'net.pnnw')
    W.save_flat(flat, width, is_fc, wts)
    n_params = sum(a.size for a in wts.values())
    for path in (graph, flat):
        w, fc, n, cs = E.inspect_net_file(path)
        assert (w, fc, n) == (width, is_fc, n_params)
        assert abs(cs - _checksum(wts)) <= 1e-9 * max(1., abs(cs))
    with pytest.raises(E.PnnError, match='neither a PNNW flat binary nor a frozen PNN graph'):
        bad = str(tmp_path / 'bad.bin')
        with open(bad, 'wb') as f:
            f.write(b'\x0a\xff\xff\xff\xff\x0f' + b'x' * 10)
        E.inspect_net_file(bad)


@pytest.mark.gpu
def test_engine_loads_frozen_graph(engine, tmp_path):
    """pnn_load_net on the frozen graph itself gives the same bits as on the exported flat binary."""
    import helpers
    for width, is_fc in ((8, True), (16, False)):
        wts = W.init_weights(width, is_fc, seed=11, bias_std=0.05, gain=helpers.GAIN.get((width, is_fc), 1.))
        graph, flat = str(tmp_path / ('g%d.pbtxt' % width)), str(tmp_path / ('n%d.pnnw' % width))
        make_graph(graph, wts, width, is_fc)
        W.save_flat(flat, width, is_fc, wts)
        img = helpers.synthetic_image(96, 128, seed=5)
        rows, cols = helpers.grid_blocks(96, 128, width)
        out = []
        for path in (flat, graph):
            engine.load_net(path)
            out.append(engine.predict_image_blocks(width, is_fc, img, rows, cols)['predictions_float32'].copy())
        numpy.testing.assert_array_equal(out[0], out[1])
