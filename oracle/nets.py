"""fp32 CPU restatement of the PNN graphs (TEST INFRASTRUCTURE, see oracle/__init__.py).

Weights are a dict {tensorflow variable name: numpy float32 array} in the
reference's own layouts:
  FC    fully_connected/weights_i [in, out], fully_connected/biases_i [out]
        (reference pnn/components.py:126-166)
  conv  convolutional/branch_{above,left}/convolution_i/{weights [k,k,Cin,Cout], biases}
        (reference pnn/tfutils.py:119-132)
  merger convolutional/merger/channelwise_fully_connected_merger/{weights [C,80,16], biases [C,16]}
        (reference pnn/tfutils.py:50-58)
  tconv convolutional/merger/transpose_convolution_i/{weights [k,k,Cout,Cin], biases}
        (reference pnn/tfutils.py:440-453)

Two implementations of every layer are kept: a fast one on torch-CPU (oneDNN /
MKL kernels, the same arithmetic class as TensorFlow-Eigen) and a slow,
literal loop version in numpy used only to validate the fast one on tiny cases.
"""
import math

import numpy
import torch
import torch.nn.functional as F

# reference pnn/PredictionNeuralNetwork.py:126-132
STRIDES_BRANCH = {
    4: (1, 1),
    8: (2, 1),
    16: (2, 1, 2, 1),
    32: (2, 2, 1, 2, 1),
    64: (2, 2, 2, 2, 1),
}
NB_HIDDEN_FC = 1200  # reference pnn/components.py:130
SLOPE = 0.1          # reference pnn/tfutils.py:192


def leaky_relu(x):
    """max(0.1*x, x) -- reference pnn/tfutils.py:176-192."""
    return torch.maximum(SLOPE * x, x)


def same_padding(n, k, s):
    """TensorFlow 'SAME' padding: (out, pad_before, pad_after).

    out = ceil(n/s); pad_total = max((out-1)*s + k - n, 0); before = total//2.
    (TensorFlow core/framework/common_shape_fns.cc GetWindowedOutputSizeVerbose;
    called through reference pnn/tfutils.py:134-137.)
    """
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


# ----------------------------------------------------------------------------
# fast torch-CPU layers
# ----------------------------------------------------------------------------

def conv2d_same(x_nhwc, w_tf, b, stride):
    """tf.nn.conv2d(..., padding='SAME') + bias_add -- reference pnn/tfutils.py:134-139.

    x_nhwc [B,H,W,Cin] float32 tensor, w_tf [k,k,Cin,Cout], b [Cout].
    """
    k = w_tf.shape[0]
    _, pt, pb = same_padding(x_nhwc.shape[1], k, stride)
    _, pl, pr = same_padding(x_nhwc.shape[2], k, stride)
    x = x_nhwc.permute(0, 3, 1, 2)
    x = F.pad(x, (pl, pr, pt, pb))
    y = F.conv2d(x, w_tf.permute(3, 2, 0, 1).contiguous(), b, stride=stride)
    return y.permute(0, 2, 3, 1).contiguous()


def tconv2d_same(x_nhwc, w_tf, b, stride):
    """tf.nn.conv2d_transpose(..., output [B,H*s,W*s,Cout], 'SAME') + bias_add.

    reference pnn/tfutils.py:455-462.  conv2d_transpose is the gradient of
    conv2d with the same padding, i.e. the full transposed convolution of size
    (n-1)*s+k cropped by the forward conv's pad_before at the start.
    w_tf [k,k,Cout,Cin].
    """
    k = w_tf.shape[0]
    h_out, w_out = x_nhwc.shape[1] * stride, x_nhwc.shape[2] * stride
    _, pt, _ = same_padding(h_out, k, stride)
    _, pl, _ = same_padding(w_out, k, stride)
    x = x_nhwc.permute(0, 3, 1, 2)
    # torch conv_transpose2d weight layout is [Cin, Cout, kh, kw]
    y = F.conv_transpose2d(x, w_tf.permute(3, 2, 0, 1).contiguous(), None, stride=stride)
    y = y[:, :, pt:pt + h_out, pl:pl + w_out]
    y = y + b.view(1, -1, 1, 1)
    return y.permute(0, 2, 3, 1).contiguous()


def channelwise_merger(in0, in1, w, b):
    """reference pnn/tfutils.py:8-73.  in0 [B,4,12,C], in1 [B,8,4,C], w [C,80,16], b [C,16]."""
    bsz, _, _, c = in0.shape
    cat = torch.cat([in0.reshape(bsz, -1, c), in1.reshape(bsz, -1, c)], dim=1)  # [B,80,C]
    cat = cat.permute(2, 0, 1)                                                  # [C,B,80]
    out = torch.bmm(cat, w) + b.unsqueeze(1)                                    # [C,B,16]
    side = int(round(math.sqrt(w.shape[2])))
    return out.permute(1, 2, 0).reshape(bsz, side, side, c).contiguous()


def _t(a):
    return torch.from_numpy(numpy.ascontiguousarray(a, dtype=numpy.float32))


def forward_fc(weights, flat):
    """reference pnn/components.py:103-180.  flat [B,5W^2] -> [B,W,W,1] (numpy float32)."""
    x = _t(flat)
    with torch.no_grad():
        for i in range(4):
            x = x @ _t(weights['fully_connected/weights_%d' % i]) + _t(weights['fully_connected/biases_%d' % i])
            if i != 3:
                x = leaky_relu(x)
    w = int(round(math.sqrt(x.shape[1])))
    return x.reshape(-1, w, w, 1).numpy()


def forward_conv(weights, above, left, return_intermediates=False):
    """reference pnn/components.py:10-101,182-261.

    above [B,W,3W,1], left [B,2W,W,1] -> [B,W,W,1] (numpy float32).
    """
    width = above.shape[1]
    strides = STRIDES_BRANCH[width]
    inter = {}
    with torch.no_grad():
        outs = []
        for name, x in (('above', _t(above)), ('left', _t(left))):
            for i, s in enumerate(strides):
                p = 'convolutional/branch_%s/convolution_%d/' % (name, i)
                x = leaky_relu(conv2d_same(x, _t(weights[p + 'weights']), _t(weights[p + 'biases']), s))
                inter['%s_%d' % (name, i)] = x
            outs.append(x)
        p = 'convolutional/merger/channelwise_fully_connected_merger/'
        x = leaky_relu(channelwise_merger(outs[0], outs[1], _t(weights[p + 'weights']), _t(weights[p + 'biases'])))
        inter['merger'] = x
        strides_m = strides[::-1]
        for i, s in enumerate(strides_m):
            p = 'convolutional/merger/transpose_convolution_%d/' % i
            x = tconv2d_same(x, _t(weights[p + 'weights']), _t(weights[p + 'biases']), s)
            if i != len(strides_m) - 1:
                x = leaky_relu(x)
            inter['tconv_%d' % i] = x
    if return_intermediates:
        return x.numpy(), {k: v.numpy() for k, v in inter.items()}
    return x.numpy()


def forward(weights, width, is_fc, inputs):
    """inputs: (flat,) for FC or (above, left) for conv."""
    if is_fc:
        return forward_fc(weights, inputs[0])
    return forward_conv(weights, inputs[0], inputs[1])


# ----------------------------------------------------------------------------
# slow literal loop versions (validation of the fast path on tiny cases only)
# ----------------------------------------------------------------------------

def conv2d_same_loops(x, w, b, s, out_dtype=numpy.float32):
    bsz, h, wd, cin = x.shape
    k, _, _, cout = w.shape
    ho, pt, _ = same_padding(h, k, s)
    wo, pl, _ = same_padding(wd, k, s)
    y = numpy.zeros((bsz, ho, wo, cout), dtype=numpy.float64)
    for oy in range(ho):
        for ox in range(wo):
            for ky in range(k):
                iy = oy * s + ky - pt
                if iy < 0 or iy >= h:
                    continue
                for kx in range(k):
                    ix = ox * s + kx - pl
                    if ix < 0 or ix >= wd:
                        continue
                    y[:, oy, ox, :] += x[:, iy, ix, :].astype(numpy.float64) @ w[ky, kx].astype(numpy.float64)
    return (y + b).astype(out_dtype)


def tconv2d_same_loops(x, w, b, s, out_dtype=numpy.float32):
    """Scatter form: every input pixel adds in*W[ky,kx] at y = iy*s + ky - pad_before."""
    bsz, h, wd, cin = x.shape
    k, _, cout, _ = w.shape
    ho, wo = h * s, wd * s
    _, pt, _ = same_padding(ho, k, s)
    _, pl, _ = same_padding(wo, k, s)
    y = numpy.zeros((bsz, ho, wo, cout), dtype=numpy.float64)
    for iy in range(h):
        for ix in range(wd):
            for ky in range(k):
                oy = iy * s + ky - pt
                if oy < 0 or oy >= ho:
                    continue
                for kx in range(k):
                    ox = ix * s + kx - pl
                    if ox < 0 or ox >= wo:
                        continue
                    # w[ky,kx] is [Cout,Cin]
                    y[:, oy, ox, :] += x[:, iy, ix, :].astype(numpy.float64) @ w[ky, kx].astype(numpy.float64).T
    return (y + b).astype(out_dtype)


def merger_loops(in0, in1, w, b, out_dtype=numpy.float32):
    bsz, _, _, c = in0.shape
    out = numpy.zeros((bsz, 16, c), dtype=numpy.float64)
    for ch in range(c):
        cat = numpy.concatenate([in0[:, :, :, ch].reshape(bsz, -1), in1[:, :, :, ch].reshape(bsz, -1)], axis=1)
        out[:, :, ch] = cat.astype(numpy.float64) @ w[ch].astype(numpy.float64) + b[ch]
    return out.reshape(bsz, 4, 4, c).astype(out_dtype)


def forward_conv_loops_float64(weights, above, left):
    """The whole convolutional PNN with the literal loop layers above, float64 from end to end: an implementation that
    shares no code with the torch path (no library convolution, no padding helper), used to make the committed goldens
    of the two pretrained checkpoints (tests/golden/make_loop_goldens.py)."""
    width = above.shape[1]
    strides = STRIDES_BRANCH[width]
    lrelu = lambda v: numpy.maximum(SLOPE * v, v)
    outs = []
    for name, x in (('above', above.astype(numpy.float64)), ('left', left.astype(numpy.float64))):
        for i, s in enumerate(strides):
            p = 'convolutional/branch_%s/convolution_%d/' % (name, i)
            x = lrelu(conv2d_same_loops(x, weights[p + 'weights'], weights[p + 'biases'], s, numpy.float64))
        outs.append(x)
    p = 'convolutional/merger/channelwise_fully_connected_merger/'
    x = lrelu(merger_loops(outs[0], outs[1], weights[p + 'weights'], weights[p + 'biases'], numpy.float64))
    strides_m = strides[::-1]
    for i, s in enumerate(strides_m):
        p = 'convolutional/merger/transpose_convolution_%d/' % i
        x = tconv2d_same_loops(x, weights[p + 'weights'], weights[p + 'biases'], s, numpy.float64)
        if i != len(strides_m) - 1:
            x = lrelu(x)
    return x


# ----------------------------------------------------------------------------
# topology helpers shared by tests / bench (shapes and work counts)
# ----------------------------------------------------------------------------

def layer_table(width, is_fc):
    """List of (kind, params...) describing the net; used for MAC counts."""
    if is_fc:
        dims = [5 * width * width, NB_HIDDEN_FC, NB_HIDDEN_FC, NB_HIDDEN_FC, width * width]
        return [('fc', dims[i], dims[i + 1]) for i in range(4)]
    strides = STRIDES_BRANCH[width]
    layers = []
    for (h, w) in ((width, 3 * width), (2 * width, width)):
        c_in, c = 1, 32
        for s in strides:
            c *= s
            h, w = h // s, w // s
            layers.append(('conv', 2 * s + 1, s, c_in, c, h * w))
            c_in = c
    layers.append(('merger', c, 80, 16))
    h = w = 4
    strides_m = strides[::-1]
    for i, s in enumerate(strides_m):
        c_out = 1 if i == len(strides_m) - 1 else c // s
        layers.append(('tconv', 2 * s + 1, s, c, c_out, h * w))  # in-pixels
        h, w, c = h * s, w * s, c_out
    return layers


def macs_per_prediction(width, is_fc):
    total = 0
    for l in layer_table(width, is_fc):
        if l[0] == 'fc':
            total += l[1] * l[2]
        elif l[0] == 'conv':
            total += l[5] * l[1] * l[1] * l[3] * l[4]
        elif l[0] == 'merger':
            total += l[1] * l[2] * l[3]
        else:
            total += l[5] * l[1] * l[1] * l[3] * l[4]
    return total
