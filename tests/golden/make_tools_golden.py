"""Golden values of the reference's own epilogue and metric functions, imported unmodified:
`tools.tools.cast_float_to_uint8` (tools/tools.py:17-49), `tools.tools.compute_psnr` (:364-401) and
`compute_performance_neural_network_vs_hevc_best_mode` (comparing_pnn_ipfcns_hevc_best_mode.py:39-88).

Run HERE (container with /root/reference): python tests/golden/make_tools_golden.py
The reference modules import matplotlib / PyQt5 / PIL / tensorflow / caffe wrappers at the top; none of them is used by
these functions and they are stubbed.
"""
import os
import sys
import types

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = ('matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker', 'PyQt5', 'PIL', 'PIL.Image', 'tensorflow', 'ipfcns', 'ipfcns.ipfcns',
         'hevc', 'hevc.intraprediction', 'hevc.intraprediction.intraprediction', 'parsing', 'parsing.parsing', 'pnn', 'pnn.batching',
         'pnn.PredictionNeuralNetwork', 'pnn.visualization', 'sets', 'sets.common')
for name in STUBS:
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules['matplotlib'].use = lambda *a, **k: None
if not hasattr(numpy, 'float'):
    numpy.float = numpy.floating     # the alias the reference was written against (removed in numpy 1.24); with the numpy of its day issubdtype(float32, numpy.float) held, which numpy.floating reproduces
sys.path.insert(0, '/root/reference')
import tools.tools as tls   # noqa: E402
import comparing_pnn_ipfcns_hevc_best_mode as cmp   # noqa: E402

rng = numpy.random.default_rng(11)
# cast: ties at .5, values outside [0, 255], float32 and float64
floats = numpy.concatenate([rng.uniform(-30., 290., 4000), numpy.arange(-3., 260., 0.5), [254.5, 255.49, 255.5, -0.5, 0.5, 1.5, 2.5]])
cast64 = tls.cast_float_to_uint8(floats.astype(numpy.float64))
cast32 = tls.cast_float_to_uint8(floats.astype(numpy.float32))
# psnr: random pairs, identical pairs (the 1e-6 guard of older versions or infinity: whatever the reference does), one-pixel differences
pairs_a = rng.integers(0, 256, (40, 8, 8), dtype=numpy.uint8)
pairs_b = pairs_a.copy()
pairs_b[:30] = numpy.clip(pairs_a[:30].astype(int) + rng.integers(-20, 21, (30, 8, 8)), 0, 255).astype(numpy.uint8)
pairs_b[30:35, 0, 0] ^= 1
psnrs = []
for a, b in zip(pairs_a, pairs_b):
    try:
        psnrs.append(float(tls.compute_psnr(a, b)))
    except Exception as e:                                   # identical arrays may be refused
        psnrs.append(float('nan'))
# win frequency
targets = rng.integers(0, 256, (200, 4, 4, 1), dtype=numpy.uint8)
preds = numpy.clip(targets.astype(int) + rng.integers(-12, 13, targets.shape), 0, 255).astype(numpy.uint8)
preds[:3] = numpy.clip(targets[:3].astype(int) + 1, 0, 255).astype(numpy.uint8)
base = rng.uniform(20., 45., 200)
psnrs_nn, frequency = cmp.compute_performance_neural_network_vs_hevc_best_mode(targets, preds, base)
base[5] = psnrs_nn[5]                                        # a tie is not a win
psnrs_nn2, frequency_tie = cmp.compute_performance_neural_network_vs_hevc_best_mode(targets, preds, base)
numpy.savez_compressed(os.path.join(HERE, 'tools_ref.npz'), floats=floats, cast64=cast64, cast32=cast32, pairs_a=pairs_a, pairs_b=pairs_b,
                       psnrs=numpy.array(psnrs), targets=targets, preds=preds, base=base, psnrs_nn=psnrs_nn2,
                       frequency=numpy.array([frequency, frequency_tie]))
print('wrote tools_ref.npz:', len(floats), 'casts,', len(psnrs), 'psnrs (nan = refused: %d),' % int(numpy.isnan(psnrs).sum()), 'frequency', frequency, frequency_tie)
