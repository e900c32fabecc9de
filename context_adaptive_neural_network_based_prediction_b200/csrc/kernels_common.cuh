// Device helpers shared by the kernels of libpnn_cuda.
#pragma once

#include "pnn_internal.h"

namespace pnn {

__device__ __forceinline__ float leaky_relu(float x) {
    // reference pnn/tfutils.py:192: tf.maximum(0.1*input, input)
    return fmaxf(0.1f * x, x);
}

// value -> (hi, lo) bf16 with hi + lo ~= value to 16 mantissa bits
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Two values at once -> packed (hi0 | hi1 << 16), (lo0 | lo1 << 16); same bits as split_bf16 on each value.  One
// cvt.rn.bf16x2.f32 (F2FP, ALU pipe) per pair and plane instead of two scalar F2F.BF16.F32, which issue at a fraction of
// the rate and made the epilogues of the small-K layers conversion-bound.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

template <bool SPLIT>
__device__ __forceinline__ float act_load(const Act& a, int64_t idx) {
    if (SPLIT) {
        return __bfloat162float(((const __nv_bfloat16*)a.p0)[idx]) + __bfloat162float(((const __nv_bfloat16*)a.p1)[idx]);
    }
    return ((const float*)a.p0)[idx];
}

template <bool SPLIT>
__device__ __forceinline__ void act_store(const Act& a, int64_t idx, float v) {
    if (SPLIT) {
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        ((__nv_bfloat16*)a.p0)[idx] = hi;
        ((__nv_bfloat16*)a.p1)[idx] = lo;
    } else {
        ((float*)a.p0)[idx] = v;
    }
}

// HM gather (reference extraction_context.cpp:56-205): element e of the flattened (above, left) context of width W from
// its raw reconstruction pixel and the availability description of pnn_internal.h (GatherHmLaunch):
//   above columns [0, W)             always copied (extraction_context.cpp:119-127)
//   above columns W + i*unit_w ...   copied iff above unit i is available (:149-166)
//   left rows [0, left_rows)         copied; the reference writer only advances on available units (:189-205)
__device__ __forceinline__ float hm_value_from_raw(int raw, int e, int W, float mean, uint32_t lo, uint32_t hi, int unit_w,
                                                   int left_rows) {
    float v = (float)raw - mean;
    const int na = 3 * W * W;
    if (e < na) {
        const int cc = e % (3 * W);
        if (cc >= W) {
            const int u = (cc - W) / unit_w;
            const uint32_t bit = u < 32 ? (lo >> u) & 1u : (hi >> (u - 32)) & 1u;
            if (!bit) v = 0.f;
        }
    } else if ((e - na) / W >= left_rows) {
        v = 0.f;
    }
    return v;
}
// the same from the staged buffer (header + pixels); unit width 0 = float bits of an already pre-processed context
__device__ __forceinline__ float hm_context_value(const int32_t* __restrict__ staged, int W, float mean, int e) {
    if (staged[2] == 0) return __int_as_float(staged[HM_HEADER_INTS + e]);
    return hm_value_from_raw(staged[HM_HEADER_INTS + e], e, W, mean, (uint32_t)staged[0], (uint32_t)staged[1], staged[2], staged[3]);
}

// Fused epilogue of the last layer: raw output, and round(clip(p + mean, 0, 255)).
//   half-even: numpy.round, reference tools/tools.py:49
//   half-away: std::round, reference TComPrediction.cpp(substitution):632
__device__ __forceinline__ void final_store(const FinalOut& f, int64_t idx, float p) {
    if (f.raw) f.raw[idx] = p;
    if (f.u8 || f.i32) {
        float v = fminf(fmaxf(p + f.mean, 0.f), 255.f);
        float r = f.round_mode == 0 ? rintf(v) : roundf(v);
        if (f.u8) f.u8[idx] = (uint8_t)r;
        if (f.i32) f.i32[idx] = (int32_t)r;
    }
}

}  // namespace pnn
