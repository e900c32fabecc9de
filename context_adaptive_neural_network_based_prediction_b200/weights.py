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
