"""Weights of the prediction neural networks: flat binary format, seeded initialisers, checkpoint export.

The engine loads weights from a flat binary ("PNNW") instead of a TensorFlow session
(replaces reference freezing_graph_pnn.py:100-143 + integration_prediction_neural_network.cpp:29-69).

Flat binary layout (little endian):
    char[8]  magic "PNNWv001"
    u32 width_target, u32 is_fully_connected, u32 n_tensors, u32 reserved
    n_tensors x { u32 name_len, char name[name_len], u32 rank, u32 dims[rank], u64 offset, u64 nbytes }
    payload: float32 tensors in the REFERENCE layouts, each 256-byte aligned, `offset` from file start.
Tensor names are the TensorFlow variable names of the reference graph:
    fully_connected/{weights,biases}_{0..3}
    convolutional/branch_{above,left}/convolution_i/{weights,biases}
    convolutional/merger/channelwise_fully_connected_merger/{weights,biases}
    convolutional/merger/transpose_convolution_i/{weights,biases}
"""
import math
import struct

import numpy

MAGIC = b'PNNWv001'

# reference pnn/PredictionNeuralNetwork.py:126-132
STRIDES_BRANCH = {4: (1, 1), 8: (2, 1), 16: (2, 1, 2, 1), 32: (2, 2, 1, 2, 1), 64: (2, 2, 2, 2, 1)}
NB_HIDDEN_FC = 1200


def tensor_shapes(width_target, is_fully_connected):
    """Ordered {name: shape} of the trainable variables of one PNN (reference pnn/components.py)."""
    shapes = {}
    if is_fully_connected:
        dims = [5 * width_target ** 2, NB_HIDDEN_FC, NB_HIDDEN_FC, NB_HIDDEN_FC, width_target ** 2]
        for i in range(4):
            shapes['fully_connected/weights_%d' % i] = (dims[i], dims[i + 1])
            shapes['fully_connected/biases_%d' % i] = (dims[i + 1],)
        return shapes
    strides = STRIDES_BRANCH[width_target]
    for name in ('above', 'left'):
        c_in, c = 1, 32
        for i, s in enumerate(strides):
            c *= s                                               # components.py:33-36
            k = 2 * s + 1                                        # tfutils.py:107
            p = 'convolutional/branch_%s/convolution_%d/' % (name, i)
            shapes[p + 'weights'] = (k, k, c_in, c)
            shapes[p + 'biases'] = (c,)
            c_in = c
    p = 'convolutional/merger/channelwise_fully_connected_merger/'
    shapes[p + 'weights'] = (c, 80, 16)                          # tfutils.py:50-58
    shapes[p + 'biases'] = (c, 16)
    strides_m = strides[::-1]
    for i, s in enumerate(strides_m):
        c_out = 1 if i == len(strides_m) - 1 else c // s         # components.py:237-244
        k = 2 * s + 1
        p = 'convolutional/merger/transpose_convolution_%d/' % i
        shapes[p + 'weights'] = (k, k, c_out, c)                 # tfutils.py:447
        shapes[p + 'biases'] = (c_out,)
        c = c_out
    return shapes


def init_weights(width_target, is_fully_connected, seed, bias_std=0., gain=1.):
    """Seeded random initialisation following the reference initialisers.

    FC: stddev 0.01 / 0.029 / 0.029 / 0.01 (reference pnn/components.py:128-166).
    Conv / tconv: 0.01 when the layer touches the pixel space, else 1/sqrt(Cin*k^2)
    (reference pnn/tfutils.py:113-117, 434-438); merger 1/sqrt(80) (tfutils.py:54).
    Biases are zero in the reference; `bias_std` > 0 draws them instead so that tests
    exercise the bias path, and `gain` scales every weight tensor (tests use it to get
    O(100) outputs like the trained nets).
    """
    rng = numpy.random.default_rng(seed)
    out = {}
    shapes = tensor_shapes(width_target, is_fully_connected)
    last_tconv = len(STRIDES_BRANCH.get(width_target, ())) - 1
    for name, shape in shapes.items():
        if name.endswith('biases') or 'biases_' in name:
            out[name] = (bias_std * rng.standard_normal(shape)).astype(numpy.float32)
            continue
        if is_fully_connected:
            std = (0.01, 0.029, 0.029, 0.01)[int(name[-1])]
        elif 'channelwise' in name:
            std = 1. / math.sqrt(80.)
        elif '/convolution_' in name:
            k, _, c_in, _ = shape
            std = 0.01 if name.split('/convolution_')[1].startswith('0/') else 1. / math.sqrt(c_in * k * k)
        else:
            k, _, _, c_in = shape
            idx = int(name.split('transpose_convolution_')[1].split('/')[0])
            std = 0.01 if idx == last_tconv else 1. / math.sqrt(c_in * k * k)
        out[name] = (gain * std * rng.standard_normal(shape)).astype(numpy.float32)
    return out


def save_flat(path, width_target, is_fully_connected, weights):
    shapes = tensor_shapes(width_target, is_fully_connected)
    names = list(shapes.keys())
    for n in names:
        if tuple(weights[n].shape) != tuple(shapes[n]):
            raise ValueError('tensor %s has shape %s, expected %s' % (n, weights[n].shape, shapes[n]))
    table_size = 8 + 16
    for n in names:
        table_size += 4 + len(n.encode()) + 4 + 4 * len(shapes[n]) + 16
    offset = (table_size + 255) // 256 * 256
    entries = []
    for n in names:
        nbytes = int(numpy.prod(shapes[n])) * 4
        entries.append((n, offset, nbytes))
        offset = (offset + nbytes + 255) // 256 * 256
    with open(path, 'wb') as f:
        f.write(MAGIC)
        f.write(struct.pack('<4I', width_target, 1 if is_fully_connected else 0, len(names), 0))
        for n, off, nbytes in entries:
            b = n.encode()
            f.write(struct.pack('<I', len(b)))
            f.write(b)
            f.write(struct.pack('<I', len(shapes[n])))
            f.write(struct.pack('<%dI' % len(shapes[n]), *shapes[n]))
            f.write(struct.pack('<2Q', off, nbytes))
        for n, off, nbytes in entries:
            f.seek(off)
            f.write(numpy.ascontiguousarray(weights[n], dtype='<f4').tobytes())
        f.seek(0, 2)
        if f.tell() < offset:
            f.write(b'\0' * (offset - f.tell()))


def load_flat(path):
    """-> (width_target, is_fully_connected, {name: float32 array})."""
    with open(path, 'rb') as f:
        data = f.read()
    if data[:8] != MAGIC:
        raise ValueError('%s is not a PNNW flat binary' % path)
    width, is_fc, n, _ = struct.unpack_from('<4I', data, 8)
    pos = 24
    out = {}
    for _ in range(n):
        (ln,) = struct.unpack_from('<I', data, pos)
        pos += 4
        name = data[pos:pos + ln].decode()
        pos += ln
        (rank,) = struct.unpack_from('<I', data, pos)
        pos += 4
        dims = struct.unpack_from('<%dI' % rank, data, pos)
        pos += 4 * rank
        off, nbytes = struct.unpack_from('<2Q', data, pos)
        pos += 16
        out[name] = numpy.frombuffer(data, dtype='<f4', count=nbytes // 4, offset=off).reshape(dims).copy()
    return width, bool(is_fc), out


# ----------------------------------------------------------------------------
# TensorFlow V2 checkpoint bundle reader (no TensorFlow needed) -- SURVEY.md appendix A
# ----------------------------------------------------------------------------

def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7f) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _block_entries(buf, offset, size):
    """Entries of one LevelDB-style table block (prefix-compressed keys)."""
    block = buf[offset:offset + size]
    (n_restarts,) = struct.unpack_from('<I', block, size - 4)
    end = size - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _parse_bundle_entry(value):
    """BundleEntryProto: 1 dtype, 2 shape{2 dim{1 size}}, 3 shard, 4 offset, 5 size, 6 crc32c."""
    pos, dtype, dims, offset, size = 0, 0, [], 0, 0
    while pos < len(value):
        tag, pos = _varint(value, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(value, pos)
            if field == 1:
                dtype = v
            elif field == 4:
                offset = v
            elif field == 5:
                size = v
        elif wire == 5:
            pos += 4
        elif wire == 2:
            ln, pos = _varint(value, pos)
            sub = value[pos:pos + ln]
            pos += ln
            if field == 2:
                sp = 0
                while sp < len(sub):
                    t2, sp = _varint(sub, sp)
                    if t2 & 7 == 2:
                        l2, sp = _varint(sub, sp)
                        dim = sub[sp:sp + l2]
                        sp += l2
                        if t2 >> 3 == 2:
                            dp = 0
                            while dp < len(dim):
                                t3, dp = _varint(dim, dp)
                                if t3 & 7 == 0:
                                    v3, dp = _varint(dim, dp)
                                    if t3 >> 3 == 1:
                                        dims.append(v3)
                                else:
                                    l3, dp = _varint(dim, dp)
                                    dp += l3
                    else:
                        _, sp = _varint(sub, sp)
        else:
            raise ValueError('unexpected wire type %d' % wire)
    return dtype, tuple(dims), offset, size


def read_tf_v2_bundle(prefix):
    """{variable name: numpy array} from `<prefix>.index` + `<prefix>.data-00000-of-00001`."""
    with open(prefix + '.index', 'rb') as f:
        idx = f.read()
    with open(prefix + '.data-00000-of-00001', 'rb') as f:
        data = f.read()
    footer = idx[-48:]
    _, p = _varint(footer, 0)          # metaindex handle
    _, p = _varint(footer, p)
    index_off, p = _varint(footer, p)
    index_size, p = _varint(footer, p)
    out = {}
    for _, handle in _block_entries(idx, index_off, index_size):
        off, hp = _varint(handle, 0)
        size, _ = _varint(handle, hp)
        for key, value in _block_entries(idx, off, size):
            if not key:
                continue                                   # bundle header
            dtype, dims, offset, size_b = _parse_bundle_entry(value)
            np_dtype = {1: '<f4', 3: '<i4'}.get(dtype)
            if np_dtype is None:
                continue
            out[key.decode()] = numpy.frombuffer(data, dtype=np_dtype, count=size_b // 4,
                                                 offset=offset).reshape(dims).copy()
    return out


def export_checkpoint(prefix, width_target, is_fully_connected, path_out):
    """TensorFlow V2 checkpoint -> flat binary (skips Adam slots and the step counters)."""
    variables = read_tf_v2_bundle(prefix)
    weights = {n: variables[n] for n in tensor_shapes(width_target, is_fully_connected)}
    save_flat(path_out, width_target, is_fully_connected, weights)
    return weights


# ----------------------------------------------------------------------------
# Frozen-graph reader (binary GraphDef written by reference freezing_graph_pnn.py:129-139; no TensorFlow needed)
# ----------------------------------------------------------------------------

def _proto_fields(buf):
    """(field number, wire type, value) of one serialized protobuf message; length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            value, pos = _varint(buf, pos)
        elif wire == 1:
            value, pos = buf[pos:pos + 8], pos + 8
        elif wire == 2:
            ln, pos = _varint(buf, pos)
            value, pos = buf[pos:pos + ln], pos + ln
        elif wire == 5:
            value, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wire)
        if pos > n:
            raise ValueError('truncated protobuf message')
        yield field, wire, value


def _parse_tensor_proto(buf):
    """TensorProto: 1 dtype, 2 tensor_shape{2 dim{1 size}}, 4 tensor_content, 5 float_val.  float32 only (DT_FLOAT = 1)."""
    dtype, dims, content, float_vals = 0, [], None, []
    for field, wire, value in _proto_fields(buf):
        if field == 1 and wire == 0:
            dtype = value
        elif field == 2 and wire == 2:
            for f2, w2, dim in _proto_fields(value):
                if f2 == 2 and w2 == 2:
                    size = 0
                    for f3, w3, v3 in _proto_fields(dim):
                        if f3 == 1 and w3 == 0:
                            size = v3
                    dims.append(size)
        elif field == 4 and wire == 2:
            content = bytes(value)
        elif field == 5:
            if wire == 2:                                   # packed
                float_vals.extend(numpy.frombuffer(bytes(value), dtype='<f4').tolist())
            elif wire == 5:
                float_vals.append(struct.unpack('<f', bytes(value))[0])
    if dtype != 1:
        return None
    count = int(numpy.prod(dims)) if dims else 1
    if content:
        arr = numpy.frombuffer(content, dtype='<f4')
        if arr.size != count:
            raise ValueError('tensor_content holds %d floats, the shape %s needs %d' % (arr.size, dims, count))
    elif len(float_vals) == count:
        arr = numpy.asarray(float_vals, dtype=numpy.float32)
    elif len(float_vals) == 1:                              # TensorFlow stores a constant-filled tensor as one value
        arr = numpy.full(count, float_vals[0], dtype=numpy.float32)
    elif not float_vals:
        arr = numpy.zeros(count, dtype=numpy.float32)
    else:
        raise ValueError('float_val holds %d values for shape %s' % (len(float_vals), dims))
    return arr.astype(numpy.float32).reshape(dims)


def read_frozen_graph(path):
    """{node name: float32 array} of the `Const` nodes of a frozen GraphDef, plus the list of (name, op) of all nodes.

    `freeze_graph` (reference freezing_graph_pnn.py:129-139) turns every variable into a `Const` node carrying the
    variable's name, so the keys are the names of `tensor_shapes`.  The file the reference calls `graph_output.pbtxt`
    is a BINARY GraphDef despite its suffix; HM reads it with `ReadBinaryProto` (integration_...cpp:29-69).
    """
    with open(path, 'rb') as f:
        data = f.read()
    if data[:64].lstrip()[:4] in (b'node', b'vers', b'libr') or data[:1] == b'#':
        raise ValueError('%s is a text-format GraphDef; the engine reads the binary file freeze_graph writes' % path)
    consts, nodes = {}, []
    view = memoryview(data)
    for field, wire, node in _proto_fields(view):
        if field != 1 or wire != 2:
            continue                                        # versions, library
        name, op, tensor = '', '', None
        for f2, w2, v2 in _proto_fields(node):
            if f2 == 1 and w2 == 2:
                name = bytes(v2).decode()
            elif f2 == 2 and w2 == 2:
                op = bytes(v2).decode()
            elif f2 == 5 and w2 == 2:                       # map<string, AttrValue> entry
                key, attr = b'', None
                for f3, w3, v3 in _proto_fields(v2):
                    if f3 == 1 and w3 == 2:
                        key = bytes(v3)
                    elif f3 == 2 and w3 == 2:
                        attr = v3
                if key == b'value' and attr is not None:
                    for f4, w4, v4 in _proto_fields(attr):
                        if f4 == 8 and w4 == 2:             # AttrValue.tensor
                            tensor = v4
        nodes.append((name, op))
        if op == 'Const' and tensor is not None:
            arr = _parse_tensor_proto(tensor)
            if arr is not None:
                consts[name] = arr
    return consts, nodes


def export_frozen_graph(path_graph, width_target, is_fully_connected, path_out):
    """Frozen graph (`graph_output.pbtxt` of the reference's HM set-up) -> flat binary."""
    consts, _ = read_frozen_graph(path_graph)
    shapes = tensor_shapes(width_target, is_fully_connected)
    missing = [n for n in shapes if n not in consts]
    if missing:
        raise ValueError('%s holds no constant named %s (is it the graph of width %d, %s?)'
                         % (path_graph, missing[0], width_target, 'fully-connected' if is_fully_connected else 'convolutional'))
    weights = {}
    for n, shape in shapes.items():
        if tuple(consts[n].shape) != tuple(shape):
            raise ValueError('%s: shape %s, expected %s' % (n, consts[n].shape, shape))
        weights[n] = consts[n]
    save_flat(path_out, width_target, is_fully_connected, weights)
    return weights


def main(argv=None):
    """python -m <pkg>.weights (--checkpoint PREFIX | --frozen-graph FILE) --width W (--fc | --conv) --out FILE.pnnw"""
    import argparse
    ap = argparse.ArgumentParser(description='export the weights of one PNN to the flat binary libpnn_cuda loads')
    src = ap.add_mutually_exclusive_group(required=True)
    src.add_argument('--checkpoint', help='TensorFlow V2 checkpoint prefix, e.g. .../model_800000.ckpt')
    src.add_argument('--frozen-graph', help='binary GraphDef written by freezing_graph_pnn.py (graph_output.pbtxt)')
    ap.add_argument('--width', type=int, required=True, choices=(4, 8, 16, 32, 64))
    kind = ap.add_mutually_exclusive_group(required=True)
    kind.add_argument('--fc', action='store_true')
    kind.add_argument('--conv', action='store_true')
    ap.add_argument('--out', required=True)
    a = ap.parse_args(argv)
    if a.checkpoint:
        w = export_checkpoint(a.checkpoint, a.width, a.fc, a.out)
    else:
        w = export_frozen_graph(a.frozen_graph, a.width, a.fc, a.out)
    print('%s: %d tensors, %d parameters' % (a.out, len(w), sum(v.size for v in w.values())))


if __name__ == '__main__':
    main()
