"""Golden values of the reference's own PYTHON best-mode path, imported unmodified:
`extract_intra_pattern`, `extract_intra_patterns`, `predict_via_hevc_best_mode` and `predict_series_via_hevc_best_mode`
(hevc/intraprediction/intraprediction.py:10-292).

Run HERE (container with /root/reference, after `make -C oracle`): python tests/golden/make_hevc_python_golden.py
The module reaches the reference's C++ through a Cython extension (hevc/intraprediction/interface.pyx), which is not built
here; it is replaced by a ctypes call into the SAME C++ file compiled unmodified (oracle/_ref/libhevc_intra_ref.so), with the
checks and the output shape of interface.pyx:9-66.  matplotlib / PyQt5 / PIL (imported by tools/tools.py) are stubbed.
"""
import ctypes
import os
import sys
import types

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker', 'PyQt5', 'PIL', 'PIL.Image'):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules['matplotlib'].use = lambda *a, **k: None
if not hasattr(numpy, 'float'):
    numpy.float = numpy.floating
if not hasattr(numpy, 'int'):
    numpy.int = numpy.integer

lib = ctypes.CDLL(os.path.join(ROOT, 'oracle', '_ref', 'libhevc_intra_ref.so'))


def predict_via_hevc_mode(intra_pattern_uint8, width_target, index_mode):
    if not intra_pattern_uint8.flags.c_contiguous:
        raise ValueError('`intra_pattern_uint8` is not C-contiguous.')
    if intra_pattern_uint8.shape[2] != 1:
        raise ValueError('`intra_pattern_uint8.shape[2]` is not equal to 1.')
    out = numpy.zeros((width_target, width_target, 1), dtype=numpy.uint8)
    code = lib.ref_hevc_intraprediction(intra_pattern_uint8.shape[0], intra_pattern_uint8.shape[1], width_target,
                                        intra_pattern_uint8.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), index_mode)
    if code != 0:
        raise RuntimeError('hevc_intraprediction threw')
    return out


interface = types.ModuleType('hevc.intraprediction.interface')
interface.predict_via_hevc_mode = predict_via_hevc_mode
sys.modules['hevc.intraprediction.interface'] = interface
sys.path.insert(0, '/root/reference')
import hevc.intraprediction.intraprediction as ref   # noqa: E402
sys.modules['hevc.intraprediction'].interface = interface      # (the attribute the import statement would have set)

sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers   # noqa: E402

out = {}
case = 0
rng = numpy.random.default_rng(23)
images = numpy.stack([helpers.synthetic_image(96, 128, s) for s in range(2)])[..., None]       # [2, 96, 128, 1]
images[1, 40:72, 30:90, 0] = 93                                                                # a flat area: PSNR ties between modes
for width in (4, 8, 16, 32):
    for masks in ((0, 0), (width, 0), (0, width), (4, 4) if width > 4 else (4, 0)):
        n = 12
        row_refs = rng.integers(0, 96 - 2 * width - 1, n)
        col_refs = rng.integers(0, 128 - 2 * width - 1, n)
        if width <= 16:
            row_refs[0], col_refs[0] = 41, 31                                                  # inside the flat area
        patterns = ref.extract_intra_patterns(images[1:2], width, row_refs, col_refs, masks)
        targets = numpy.stack([images[1, r + 1:r + 1 + width, c + 1:c + 1 + width, :] for r, c in zip(row_refs, col_refs)])
        indices, psnrs, preds = ref.predict_series_via_hevc_best_mode(patterns, targets)
        out['c%d_width' % case] = numpy.array([width])
        out['c%d_masks' % case] = numpy.array(masks)
        out['c%d_row_refs' % case] = row_refs
        out['c%d_col_refs' % case] = col_refs
        out['c%d_patterns' % case] = patterns
        out['c%d_indices' % case] = numpy.asarray(indices)
        out['c%d_psnrs' % case] = numpy.asarray(psnrs)
        out['c%d_preds' % case] = preds
        case += 1
out['n_cases'] = numpy.array([case])
out['image'] = images[1, :, :, 0]
numpy.savez_compressed(os.path.join(HERE, 'hevc_python_ref.npz'), **out)
print('wrote hevc_python_ref.npz:', case, 'cases')
