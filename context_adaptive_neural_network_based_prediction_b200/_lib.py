"""ctypes binding of libpnn_cuda.so (the C ABI declared in include/pnn_cuda.h).

There is no fallback: if the shared library is missing the import fails loudly, and
`pnn_create` itself fails when no sm_100 GPU is present.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libpnn_cuda.so')

# every symbol include/pnn_cuda.h declares
EXPORTS = (
    'pnn_create', 'pnn_destroy', 'pnn_last_error', 'pnn_load_net', 'pnn_set_precision',
    'pnn_set_context', 'pnn_predict_hm', 'pnn_predict_batch', 'pnn_predict_image_blocks',
    'pnn_predict_batch_device', 'pnn_predict_image_blocks_device', 'pnn_launch_count',
    'pnn_last_hm_device_ms', 'pnn_version', 'pnn_debug_get_activation', 'pnn_win_flags_device', 'pnn_set_profiling',
    'pnn_profile_report', 'pnn_debug_time_gemm', 'pnn_predict_hm_context', 'pnn_set_hm_fused', 'pnn_hevc_best_mode', 'pnn_hevc_best_mode_device',
    'pnn_inspect_net_file', 'pnn_predict_image_blocks_async', 'pnn_synchronize',
    'pnn_set_hm_cache', 'pnn_hm_cache_stats', 'pnn_set_workspace_budget', 'pnn_register_net', 'pnn_set_context_lazy', 'pnn_create_deferred', 'pnn_release_at_exit', 'pnn_warm_up',
    'pnn_predict_hm_begin',
)

PRECISION_FP32 = 0
PRECISION_BF16X3 = 1

_lib = None


def load():
    """Loads libpnn_cuda.so once; raises if it has not been built (run `python -c "import __graft_entry__ as g; g.build()"`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('%s is missing: build it with __graft_entry__.build() (make -C %s). '
                          'There is no CPU fallback.' % (LIB_PATH, os.path.dirname(LIB_PATH)))
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i32, i64 = c.c_void_p, c.c_int, c.c_int64
    lib.pnn_create.argtypes = [c.c_char_p, c.c_float, i32, i32, c.POINTER(vp)]
    lib.pnn_create.restype = i32
    lib.pnn_create_deferred.argtypes = lib.pnn_create.argtypes
    lib.pnn_create_deferred.restype = i32
    lib.pnn_warm_up.argtypes = [vp]
    lib.pnn_warm_up.restype = i32
    lib.pnn_release_at_exit.argtypes = [vp]
    lib.pnn_release_at_exit.restype = i32
    lib.pnn_destroy.argtypes = [vp]
    lib.pnn_destroy.restype = None
    lib.pnn_last_error.argtypes = [vp]
    lib.pnn_last_error.restype = c.c_char_p
    lib.pnn_load_net.argtypes = [vp, c.c_char_p]
    lib.pnn_load_net.restype = i32
    lib.pnn_register_net.argtypes = [vp, c.c_char_p]
    lib.pnn_register_net.restype = i32
    lib.pnn_inspect_net_file.argtypes = [c.c_char_p, c.POINTER(c.c_int), c.POINTER(c.c_int), c.POINTER(c.c_int64),
                                         c.POINTER(c.c_double)]
    lib.pnn_inspect_net_file.restype = i32
    lib.pnn_set_precision.argtypes = [vp, i32]
    lib.pnn_set_precision.restype = i32
    lib.pnn_set_context.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32]
    lib.pnn_set_context.restype = i32
    lib.pnn_predict_hm.argtypes = [vp, i32, vp, i32]
    lib.pnn_predict_hm.restype = i32
    lib.pnn_predict_hm_begin.argtypes = [vp, i32]
    lib.pnn_predict_hm_begin.restype = i32
    lib.pnn_predict_batch.argtypes = [vp, i32, i32, vp, vp, i64, vp]
    lib.pnn_predict_batch.restype = i32
    lib.pnn_predict_image_blocks.argtypes = [vp, i32, i32, vp, i32, i32, i32, vp, vp, vp, i64, i32, i32, vp, vp, vp]
    lib.pnn_predict_image_blocks.restype = i32
    lib.pnn_predict_image_blocks_async.argtypes = lib.pnn_predict_image_blocks.argtypes
    lib.pnn_predict_image_blocks_async.restype = i32
    lib.pnn_synchronize.argtypes = [vp]
    lib.pnn_synchronize.restype = i32
    lib.pnn_predict_batch_device.argtypes = [vp, i32, i32, vp, vp, i64, vp, vp]
    lib.pnn_predict_batch_device.restype = i32
    lib.pnn_predict_image_blocks_device.argtypes = [vp, i32, i32, vp, i32, i32, i32, vp, vp, vp, i64, i32, i32,
                                                    vp, vp, vp, vp]
    lib.pnn_predict_image_blocks_device.restype = i32
    lib.pnn_launch_count.argtypes = [vp]
    lib.pnn_launch_count.restype = i64
    lib.pnn_last_hm_device_ms.argtypes = [vp]
    lib.pnn_last_hm_device_ms.restype = c.c_float
    lib.pnn_debug_get_activation.argtypes = [vp, i32, i32, i32, i64, vp, c.POINTER(i64)]
    lib.pnn_debug_get_activation.restype = i32
    lib.pnn_win_flags_device.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.pnn_win_flags_device.restype = i32
    lib.pnn_set_profiling.argtypes = [vp, i32]
    lib.pnn_set_profiling.restype = i32
    lib.pnn_profile_report.argtypes = [vp, c.POINTER(c.c_double), c.POINTER(c.c_double), c.POINTER(i64), c.POINTER(c.c_double)]
    lib.pnn_profile_report.restype = c.c_char_p
    lib.pnn_debug_time_gemm.argtypes = [vp, i64, i32, i32, i32, i32]
    lib.pnn_debug_time_gemm.restype = c.c_float
    lib.pnn_predict_hm_context.argtypes = [vp, i32, vp, vp, vp]
    lib.pnn_predict_hm_context.restype = i32
    lib.pnn_set_hm_fused.argtypes = [vp, i32]
    lib.pnn_set_hm_fused.restype = i32
    lib.pnn_hevc_best_mode.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp, vp, i64, i32, i32, vp, vp, vp]
    lib.pnn_hevc_best_mode.restype = i32
    lib.pnn_hevc_best_mode_device.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp, vp, i64, i32, i32, vp, vp, vp, vp]
    lib.pnn_hevc_best_mode_device.restype = i32
    lib.pnn_set_context_lazy.argtypes = [vp, i32]
    lib.pnn_set_context_lazy.restype = i32
    lib.pnn_set_hm_cache.argtypes = [vp, i32]
    lib.pnn_set_hm_cache.restype = i32
    lib.pnn_hm_cache_stats.argtypes = [vp, c.POINTER(i64), c.POINTER(i64)]
    lib.pnn_hm_cache_stats.restype = i32
    lib.pnn_set_workspace_budget.argtypes = [vp, i64]
    lib.pnn_set_workspace_budget.restype = i32
    lib.pnn_version.argtypes = []
    lib.pnn_version.restype = c.c_char_p
    _lib = lib
    return lib
