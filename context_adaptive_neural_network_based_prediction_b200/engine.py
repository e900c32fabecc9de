"""Python host side over the C ABI of libpnn_cuda (numpy in, numpy out; device-pointer variants for torch)."""
import ctypes

import numpy

from . import _lib

MEAN_TRAINING_LUMINANCE = 117.8952234192841   # reference sets/results/training_set/means/luminance/mean_training.pkl


class PnnError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def inspect_net_file(path):
    """Host-only: (width_target, is_fully_connected, n_parameters, checksum) of a PNNW file or a frozen graph, parsed
    by the library exactly as `Engine.load_net` parses it (C ABI pnn_inspect_net_file).  Needs no GPU."""
    lib = _lib.load()
    w, fc, n, cs = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_double()
    if lib.pnn_inspect_net_file(path.encode(), ctypes.byref(w), ctypes.byref(fc), ctypes.byref(n), ctypes.byref(cs)) != 0:
        raise PnnError(lib.pnn_last_error(None).decode())
    return w.value, bool(fc.value), n.value, cs.value


class Engine(object):
    """One `pnn_handle`: the nets loaded on one GPU.

    Replaces the TensorFlow session(s) of the reference: `tf.Session` + `Saver.restore` on the
    Python side (pnn/PredictionNeuralNetwork.py:185-200) and the five `tensorflow::Session`s of
    HM (TComPrediction.cpp:108-236).
    """

    def __init__(self, mean_training=MEAN_TRAINING_LUMINANCE, device=0, paths_file=None, qp_selection=22, deferred=False):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        self._pending = []                      # arrays of enqueued asynchronous calls, kept alive until synchronize()
        self.mean_training = float(mean_training)
        path = paths_file.encode() if paths_file else None
        create = self._lib.pnn_create_deferred if deferred else self._lib.pnn_create   # deferred: the GPU is touched at first use
        if create(path, ctypes.c_float(self.mean_training), int(qp_selection), int(device), ctypes.byref(self._h)) != 0:
            raise PnnError(self._lib.pnn_last_error(None).decode())

    def warm_up(self):
        """After `deferred=True`: device initialisation and net uploads start on a thread of the library (pnn_warm_up)."""
        self._check(self._lib.pnn_warm_up(self._h))

    def close(self):
        if getattr(self, '_h', None) and self._h.value:
            self._lib.pnn_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, code):
        if code != 0:
            raise PnnError(self._lib.pnn_last_error(self._h).decode())

    # ------------------------------------------------------------------ set-up
    def load_net(self, path):
        self._check(self._lib.pnn_load_net(self._h, path.encode()))

    def register_net(self, path):
        """Validates the file now, uploads the weights at first use (C ABI pnn_register_net)."""
        self._check(self._lib.pnn_register_net(self._h, path.encode()))

    def set_precision(self, precision):
        """'fp32' (FFMA yard-stick) or 'bf16x3' (tcgen05, default)."""
        code = {'fp32': _lib.PRECISION_FP32, 'bf16x3': _lib.PRECISION_BF16X3}[precision]
        self._check(self._lib.pnn_set_precision(self._h, code))

    @property
    def launch_count(self):
        return int(self._lib.pnn_launch_count(self._h))

    @property
    def last_hm_device_ms(self):
        return float(self._lib.pnn_last_hm_device_ms(self._h))

    def set_hm_fused(self, enabled):
        """In-loop nets: persistent shared-memory-resident FC kernel + split-K conv graphs (default) or plain CUDA graphs."""
        self._check(self._lib.pnn_set_hm_fused(self._h, int(bool(enabled))))

    def set_context_lazy(self, enabled):
        """pnn_set_context records its arguments only; pnn_predict_hm copies the pixels (C ABI pnn_set_context_lazy)."""
        self._check(self._lib.pnn_set_context_lazy(self._h, int(bool(enabled))))

    def set_hm_cache(self, enabled):
        """Memo of in-loop results keyed by the exact context (C ABI pnn_set_hm_cache); off by default."""
        self._check(self._lib.pnn_set_hm_cache(self._h, int(bool(enabled))))

    @property
    def hm_cache_stats(self):
        hits, misses = ctypes.c_int64(), ctypes.c_int64()
        self._check(self._lib.pnn_hm_cache_stats(self._h, ctypes.byref(hits), ctypes.byref(misses)))
        return hits.value, misses.value

    def set_workspace_budget(self, bytes_per_net):
        """Activation workspace one net may use; batched calls are cut into chunks that fit (results do not depend on it)."""
        self._check(self._lib.pnn_set_workspace_budget(self._h, int(bytes_per_net)))

    def set_profiling(self, enabled):
        self._check(self._lib.pnn_set_profiling(self._h, int(bool(enabled))))

    def profile_report(self):
        """-> dict(text, gemm_ms, gemm_flops, gemm_launches, other_ms); resets the counters."""
        g_ms, g_fl, o_ms = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        g_n = ctypes.c_int64()
        text = self._lib.pnn_profile_report(self._h, ctypes.byref(g_ms), ctypes.byref(g_fl), ctypes.byref(g_n),
                                            ctypes.byref(o_ms)).decode()
        return {'text': text, 'gemm_ms': g_ms.value, 'gemm_flops': g_fl.value, 'gemm_launches': g_n.value,
                'other_ms': o_ms.value}

    def win_flags_device(self, d_psnr, d_baseline, n, d_win, stream=0):
        self._check(self._lib.pnn_win_flags_device(self._h, d_psnr, d_baseline, n, d_win, stream))

    def time_gemm(self, m, n, k, iters=5, flags=0):
        """Tuning aid: mean ms per launch of the tcgen05 GEMM kernel on a synthetic [m,k]x[k,n] problem."""
        ms = float(self._lib.pnn_debug_time_gemm(self._h, m, n, k, iters, flags))
        if ms < 0:
            raise PnnError(self._lib.pnn_last_error(self._h).decode())
        return ms

    def get_activation(self, width_target, is_fully_connected, buffer_index, n_samples):
        """Inspection hook: activation buffer `buffer_index` as left by the last call, float32 [n, elems]."""
        per = ctypes.c_int64()
        self._check(self._lib.pnn_debug_get_activation(self._h, width_target, int(bool(is_fully_connected)), buffer_index,
                                                       0, None, ctypes.byref(per)))
        out = numpy.empty((n_samples, per.value), dtype=numpy.float32)
        self._check(self._lib.pnn_debug_get_activation(self._h, width_target, int(bool(is_fully_connected)), buffer_index,
                                                       n_samples, _ptr(out), ctypes.byref(per)))
        return out

    # ------------------------------------------------------------------ offline path, host buffers
    def predict_batch(self, width_target, is_fully_connected, above_or_flat, left=None):
        """pnn.batching.predict_by_batch_via_pnn equivalent: float32 contexts -> float32 [N, W, W, 1]."""
        a = numpy.ascontiguousarray(above_or_flat, dtype=numpy.float32)
        n = a.shape[0]
        px = width_target * width_target
        if is_fully_connected:
            if a.size != n * 5 * px:
                raise ValueError('the flattened contexts must have shape [N, 5*W*W]')
            l = None
        else:
            l = numpy.ascontiguousarray(left, dtype=numpy.float32)
            if a.size != n * 3 * px or l.size != n * 2 * px:
                raise ValueError('the context portions must have shapes [N, W, 3W, 1] and [N, 2W, W, 1]')
        out = numpy.empty((n, width_target, width_target, 1), dtype=numpy.float32)
        self._check(self._lib.pnn_predict_batch(self._h, width_target, int(bool(is_fully_connected)), _ptr(a), _ptr(l),
                                                n, _ptr(out)))
        return out

    def predict_image_blocks(self, width_target, is_fully_connected, images_uint8, rows, cols, image_index=None,
                             masks=(0, 0), want_float=True, want_uint8=True, want_psnr=True,
                             out_float=None, out_uint8=None, out_psnr=None, wait=True):
        """Fused gather + net + epilogue for blocks of uint8 images [n_images, H, W_img] (or [H, W_img]).

        Returns a dict with 'predictions_float32' [N, W, W] (raw), 'predictions_uint8' [N, W, W] and
        'psnrs' [N] (float64) for the requested outputs.

        `wait=False` (C ABI pnn_predict_image_blocks_async) returns once the work is enqueued: uploads, kernels and
        read-backs of successive calls overlap, the outputs are complete after `synchronize()` (pass pinned arrays for
        real overlap; do not modify the inputs before `synchronize()`).
        """
        img = numpy.ascontiguousarray(images_uint8, dtype=numpy.uint8)
        if img.ndim == 2:
            img = img[None]
        rows = numpy.ascontiguousarray(rows, dtype=numpy.int32)
        cols = numpy.ascontiguousarray(cols, dtype=numpy.int32)
        n = rows.shape[0]
        idx = None if image_index is None else numpy.ascontiguousarray(image_index, dtype=numpy.int32)
        w = width_target
        # caller-provided (e.g. pinned) output arrays are used as they are
        f32 = out_float if out_float is not None else (numpy.empty((n, w, w), dtype=numpy.float32) if want_float else None)
        u8 = out_uint8 if out_uint8 is not None else (numpy.empty((n, w, w), dtype=numpy.uint8) if want_uint8 else None)
        psnr = out_psnr if out_psnr is not None else (numpy.empty((n,), dtype=numpy.float64) if want_psnr else None)
        for arr, dt, count in ((f32, numpy.float32, n * w * w), (u8, numpy.uint8, n * w * w), (psnr, numpy.float64, n)):
            if arr is not None and (arr.dtype != dt or arr.size != count or not arr.flags['C_CONTIGUOUS']):
                raise ValueError('an output array has the wrong dtype, size or layout')
        if not wait:
            # the arrays the library reads and writes after this call returns stay referenced until synchronize()
            self._pending.append((img, rows, cols, idx, f32, u8, psnr))
        call = self._lib.pnn_predict_image_blocks if wait else self._lib.pnn_predict_image_blocks_async
        self._check(call(
            self._h, w, int(bool(is_fully_connected)), _ptr(img), img.shape[0], img.shape[1], img.shape[2],
            _ptr(idx), _ptr(rows), _ptr(cols), n, int(masks[0]), int(masks[1]), _ptr(f32), _ptr(u8), _ptr(psnr)))
        out = {}
        if f32 is not None:
            out['predictions_float32'] = f32
        if u8 is not None:
            out['predictions_uint8'] = u8
        if psnr is not None:
            out['psnrs'] = psnr
        return out

    def synchronize(self):
        """Waits for the calls made with `wait=False` (C ABI pnn_synchronize)."""
        self._check(self._lib.pnn_synchronize(self._h))
        self._pending = []

    def hevc_best_mode(self, width_target, images_uint8, rows, cols, image_index=None, masks=(0, 0), want_predictions=True):
        """Best of the 35 HEVC intra modes per block (reference hevc/intraprediction/intraprediction.py:183-292).

        Returns a dict with 'indices_hevc_best_mode' uint8 [N], 'psnrs_hevc_best_mode' float64 [N] and, if requested,
        'predictions_hevc_best_mode_uint8' [N, W, W].
        """
        img = numpy.ascontiguousarray(images_uint8, dtype=numpy.uint8)
        if img.ndim == 2:
            img = img[None]
        rows = numpy.ascontiguousarray(rows, dtype=numpy.int32)
        cols = numpy.ascontiguousarray(cols, dtype=numpy.int32)
        idx = None if image_index is None else numpy.ascontiguousarray(image_index, dtype=numpy.int32)
        n, w = rows.shape[0], width_target
        best = numpy.empty(n, dtype=numpy.uint8)
        psnr = numpy.empty(n, dtype=numpy.float64)
        pred = numpy.empty((n, w, w), dtype=numpy.uint8) if want_predictions else None
        self._check(self._lib.pnn_hevc_best_mode(self._h, w, _ptr(img), img.shape[0], img.shape[1], img.shape[2], _ptr(idx),
                                                 _ptr(rows), _ptr(cols), n, int(masks[0]), int(masks[1]), _ptr(best), _ptr(psnr),
                                                 _ptr(pred)))
        out = {'indices_hevc_best_mode': best, 'psnrs_hevc_best_mode': psnr}
        if want_predictions:
            out['predictions_hevc_best_mode_uint8'] = pred
        return out

    def hevc_best_mode_device(self, width_target, d_images, n_images, height, width_image, d_image_index, d_rows, d_cols, n,
                              masks, d_best_index, d_psnr, d_pred_u8, stream=0):
        self._check(self._lib.pnn_hevc_best_mode_device(self._h, width_target, d_images, n_images, height, width_image,
                                                        d_image_index, d_rows, d_cols, n, int(masks[0]), int(masks[1]),
                                                        d_best_index, d_psnr, d_pred_u8, stream))

    # ------------------------------------------------------------------ offline path, device buffers
    def predict_image_blocks_device(self, width_target, is_fully_connected, d_images, n_images, height, width_image,
                                    d_image_index, d_rows, d_cols, n, masks, d_out_f32, d_out_u8, d_out_psnr,
                                    stream=0):
        """All pointers are integers (e.g. torch.Tensor.data_ptr()); asynchronous on `stream`."""
        self._check(self._lib.pnn_predict_image_blocks_device(
            self._h, width_target, int(bool(is_fully_connected)), d_images, n_images, height, width_image,
            d_image_index, d_rows, d_cols, n, int(masks[0]), int(masks[1]), d_out_f32, d_out_u8, d_out_psnr,
            stream))

    def predict_batch_device(self, width_target, is_fully_connected, d_above_or_flat, d_left, n, d_out, stream=0):
        self._check(self._lib.pnn_predict_batch_device(self._h, width_target, int(bool(is_fully_connected)),
                                                       d_above_or_flat, d_left, n, d_out, stream))

    # ------------------------------------------------------------------ in-loop (HM) path
    def set_context(self, width, plane_int32, origin_row, origin_col, neighbor_flags, num_intra_neighbor,
                    unit_width=4, unit_height=4, above_units=None, left_units=None):
        """Mirrors extract_context_portions (reference extraction_context.h:36-48) on a numpy int32 plane."""
        plane = plane_int32
        if plane.dtype != numpy.int32 or not plane.flags['C_CONTIGUOUS'] or plane.ndim != 2:
            raise ValueError('`plane_int32` must be a C-contiguous 2D int32 array')
        stride = plane.shape[1]
        if above_units is None:
            above_units = 2 * width // unit_width
        if left_units is None:
            left_units = 2 * width // unit_height
        flags = numpy.ascontiguousarray(neighbor_flags, dtype=numpy.uint8)
        origin = plane.ctypes.data + 4 * (origin_row * stride + origin_col)
        self._keep = (plane, flags)
        self._check(self._lib.pnn_set_context(self._h, width, ctypes.c_void_p(origin), stride, _ptr(flags),
                                              int(num_intra_neighbor), unit_width, unit_height, above_units, left_units))

    def predict_hm_context(self, width, above_or_flat, left=None):
        """Batch-1 call with an already extracted float context (what Session::Run does in the reference's HM)."""
        a = numpy.ascontiguousarray(above_or_flat, dtype=numpy.float32)
        l = None if left is None else numpy.ascontiguousarray(left, dtype=numpy.float32)
        out = numpy.empty((width, width), dtype=numpy.float32)
        self._check(self._lib.pnn_predict_hm_context(self._h, width, _ptr(a), _ptr(l), _ptr(out)))
        return out

    def predict_hm_begin(self, width):
        """Posts the context of the last `set_context` to the GPU and returns at once; `predict_hm` collects the answer."""
        self._check(self._lib.pnn_predict_hm_begin(self._h, width))

    def predict_hm(self, width, dst_stride=None):
        """NN branch of predIntraAng: returns the int32 [W, W] prediction (HM rounding)."""
        stride = width if dst_stride is None else dst_stride
        dst = numpy.zeros((width, stride), dtype=numpy.int32)
        self._check(self._lib.pnn_predict_hm(self._h, width, _ptr(dst), stride))
        return dst[:, :width]
