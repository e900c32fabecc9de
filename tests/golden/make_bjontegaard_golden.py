"""Golden values of the reference's own `compute_bjontegaard` (tools/tools.py:256-362), imported unmodified.

Run HERE (container with /root/reference): python tests/golden/make_bjontegaard_golden.py
The reference module imports matplotlib / PyQt5 / PIL at the top; they are not needed by this function and are stubbed.
"""
import os
import sys
import types

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker', 'PyQt5', 'PIL', 'PIL.Image'):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules['matplotlib'].use = lambda *a, **k: None
sys.path.insert(0, '/root/reference')
import tools.tools as tls   # noqa: E402

rng = numpy.random.default_rng(7)
cases = {}
# the reference's own test curves (test_tools.py:164-167)
r0, r1 = numpy.linspace(0.15, 2.05, num=191), numpy.linspace(0.1, 1.7, num=321)
cases['ref_test'] = (r0, 40. * numpy.sqrt(r0), r1, 20. * numpy.sqrt(r1) + 10.)
# four-point curves as an encoder produces them (QP 22 / 27 / 32 / 37)
for i in range(6):
    rates = numpy.sort(rng.uniform(0.05, 2.5, 4))[::-1].copy()
    psnrs = 30. + 6. * numpy.log2(rates / rates[-1]) + rng.normal(0., 0.05, 4)
    scale = rng.uniform(0.85, 1.15)
    cases['four_%d' % i] = (rates, psnrs, rates * scale, psnrs + rng.normal(0., 0.1, 4))
out = {}
for name, (a, b, c, d) in cases.items():
    out[name + '_rates_0'], out[name + '_psnrs_0'], out[name + '_rates_1'], out[name + '_psnrs_1'] = a, b, c, d
    out[name + '_value'] = numpy.float64(tls.compute_bjontegaard(a, b, c, d))
    print(name, out[name + '_value'])
numpy.savez_compressed(os.path.join(HERE, 'bjontegaard_ref.npz'), **out)
