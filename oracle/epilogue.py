"""Epilogues and statistics of the PNN path on the CPU (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
import numpy


def epilogue_hm(pred_f32, mean):
    """reference TComPrediction.cpp(substitution):621-635.

    static_cast<int>(std::round(max(0, min(255, p + mean)))) with float32 arithmetic and
    round-half-away-from-zero; returns int32.
    """
    v = pred_f32.astype(numpy.float32) + numpy.float32(mean)
    v = numpy.clip(v, numpy.float32(0.), numpy.float32(255.))
    # std::round on a non-negative float: floor(v + 0.5) is exact here because
    # v <= 255 has at least 16 fractional mantissa bits left (no double rounding)
    return numpy.floor(v.astype(numpy.float64) + 0.5).astype(numpy.int32)


def epilogue_numpy(pred_f32, mean):
    """reference comparing_pnn_ipfcns_hevc_best_mode.py:259 + tools/tools.py:49.

    numpy.round(clip(pred + mean, 0, 255)).astype(uint8); numpy.round is half-to-even; the
    sum stays float32 (float32 array + Python float).
    """
    v = pred_f32.astype(numpy.float32) + numpy.float32(mean)
    return numpy.round(v.clip(min=0., max=255.)).astype(numpy.uint8)


def psnr(a_u8, b_u8):
    """reference tools/tools.py:364-401: 10*log10(255^2 / (mse + 1e-6)) in float64."""
    d = a_u8.astype(numpy.float64) - b_u8.astype(numpy.float64)
    return 10. * numpy.log10(255. ** 2 / (numpy.mean(d ** 2) + 1.e-6))


def performance_vs_baseline(targets_u8, preds_u8, psnrs_baseline):
    """reference comparing_pnn_ipfcns_hevc_best_mode.py:39-88 -> (psnrs float64 [N], win frequency)."""
    n = targets_u8.shape[0]
    out = numpy.zeros(n)
    for i in range(n):
        out[i] = psnr(targets_u8[i], preds_u8[i])
    freq = float(numpy.count_nonzero(out - psnrs_baseline > 0.)) / n
    return out, freq
