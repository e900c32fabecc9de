// HEVC intra prediction baseline of the offline evaluation: the 35 modes of HM-16.15 on an unfiltered intra pattern
// and the best mode by prediction PSNR (reference hevc/intraprediction/c++/source/extracted_hevc_intraprediction.cpp:3-421,
// hevc/intraprediction/intraprediction.py:8-292).  Integer arithmetic, bit-exact.  One warp per target block: the
// reference rows live in shared memory, every lane evaluates its share of the W*W pixels for each mode, the squared
// errors are added with a shuffle tree (exact integers), the first mode with the lowest error wins (the reference keeps
// the first mode with the strictly highest PSNR, which is the same rule).
#include "kernels_common.cuh"

namespace pnn {

namespace {

constexpr int WARPS = 8;
constexpr int REF_LEN = 2 * 64 + 1;          // above / left: corner + 2W pixels
constexpr int MAIN_LEN = 3 * 64 + 2;         // projected main reference: indices -(W-1) .. 2W

__constant__ int c_ang_table[9] = {0, 2, 5, 9, 13, 17, 21, 26, 32};
__constant__ int c_inv_ang_table[9] = {0, 4096, 1638, 910, 630, 482, 390, 315, 256};

struct ModeInfo {
    bool is_ver;
    int angle, inv_angle;
};

__device__ __forceinline__ ModeInfo mode_info(int mode) {
    ModeInfo m;
    m.is_ver = mode >= 18;                                              // extracted_hevc_intraprediction.cpp:163
    const int ang_mode = m.is_ver ? mode - 26 : -(mode - 10);
    const int abs_mode = ang_mode < 0 ? -ang_mode : ang_mode;
    m.angle = (ang_mode < 0 ? -1 : 1) * c_ang_table[abs_mode];
    m.inv_angle = c_inv_ang_table[abs_mode];
    return m;
}

// prediction of pixel (y, x); ref_main already built for angular modes (index off + i holds refMain[i])
__device__ __forceinline__ int predict_pixel(int mode, const ModeInfo& mi, int y, int x, int W, int shift, int dc,
                                             const int* above, const int* left, const int* ref_main) {
    if (mode == 0) {                                                    // planar, :324-383
        const int hor = (left[y + 1] << shift) + W + (x + 1) * (above[W + 1] - left[y + 1]);
        const int ver = (above[x + 1] << shift) + (y + 1) * (left[W + 1] - above[x + 1]);
        return (hor + ver) >> (shift + 1);
    }
    if (mode == 1) {                                                    // DC + edge filter, :286-322, :385-421
        if (W <= 16) {
            if (y == 0 && x == 0) return (above[1] + left[1] + 2 * dc + 2) >> 2;
            if (y == 0) return (above[x + 1] + 3 * dc + 2) >> 2;
            if (x == 0) return (left[y + 1] + 3 * dc + 2) >> 2;
        }
        return dc;
    }
    const int off = W - 1;
    const int yy = mi.is_ver ? y : x, xx = mi.is_ver ? x : y;           // horizontal modes are computed transposed, :201-207, :268-280
    if (mi.angle == 0) {                                                // pure vertical / horizontal, :208-225
        int v = ref_main[off + xx + 1];
        if (W <= 16 && xx == 0) {
            const int* side = mi.is_ver ? left : above;
            v += (side[yy + 1] - side[0]) >> 1;
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
        }
        return v;
    }
    const int delta = (yy + 1) * mi.angle;                              // :229-256
    const int d_int = delta >> 5, d_fract = delta & 31;
    const int* p = ref_main + off + xx + d_int + 1;
    return d_fract ? ((32 - d_fract) * p[0] + d_fract * p[1] + 16) >> 5 : p[0];
}

__device__ __forceinline__ void build_ref_main(const ModeInfo& mi, int W, const int* above, const int* left, int* ref_main, int lane) {
    const int* main_src = mi.is_ver ? above : left;
    const int* side_src = mi.is_ver ? left : above;
    const int off = W - 1;
    if (mi.angle < 0) {                                                 // :173-194
        for (int i = lane; i <= W; i += 32) ref_main[off + i] = main_src[i];
        const int last = (W * mi.angle) >> 5;                           // k runs from -1 down to last + 1
        for (int k = -1 - lane; k > last; k -= 32) ref_main[off + k] = side_src[(128 + (-k) * mi.inv_angle) >> 8];
    } else {                                                            // :195-206
        for (int i = lane; i <= 2 * W; i += 32) ref_main[off + i] = main_src[i];
    }
}

struct HevcLaunch {
    const uint8_t* images;
    const int32_t* image_index;
    const int32_t* rows;
    const int32_t* cols;
    int64_t n;
    int H, Wimg, W, mask_w, mask_h, n_images;
    uint8_t* best_index;
    double* psnr;
    uint8_t* pred;
};

__global__ void __launch_bounds__(WARPS * 32) hevc_best_mode_kernel(HevcLaunch L) {
    __shared__ int s_above[WARPS][REF_LEN];
    __shared__ int s_left[WARPS][REF_LEN];
    __shared__ int s_main[WARPS][MAIN_LEN];
    __shared__ uint8_t s_target[WARPS][64 * 64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* above = s_above[warp];
    int* left = s_left[warp];
    int* ref_main = s_main[warp];
    uint8_t* target = s_target[warp];
    const int W = L.W, shift = 31 - __clz(W), px = W * W;
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    for (int64_t i = (int64_t)blockIdx.x * WARPS + warp; i < L.n; i += nwarps) {
        const int img_i = L.image_index ? L.image_index[i] : 0;
        const int r0 = L.rows[i], c0 = L.cols[i];
        // the device-pointer entry point cannot check its block list on the host: a block (or the first pixel of its intra
        // pattern) outside the image, or a bad image index, yields best_index 255 / PSNR NaN / a zero prediction instead of
        // reads out of bounds
        if (img_i < 0 || img_i >= L.n_images || r0 < 1 || c0 < 1 || r0 + W > L.H || c0 + W > L.Wimg) {
            if (lane == 0) {
                if (L.best_index) L.best_index[i] = 255;
                if (L.psnr) L.psnr[i] = __longlong_as_double(0x7ff8000000000000LL);
            }
            if (L.pred) {
                for (int p = lane; p < px; p += 32) L.pred[i * px + p] = 0;
            }
            continue;
        }
        const uint8_t* img = L.images + (int64_t)img_i * L.H * L.Wimg;
        const int rr = r0 - 1, cr = c0 - 1;                             // comparing_pnn_ipfcns_hevc_best_mode.py:234-235
        // intraprediction.py:73-88 (pattern size) + extracted_hevc_intraprediction.cpp:34-84 (padding with the last pixel)
        int wp = 2 * W + 1 - L.mask_w, hp = 2 * W + 1 - L.mask_h;
        wp = min(wp, L.Wimg - cr);
        hp = min(hp, L.H - rr);
        for (int k = lane; k <= 2 * W; k += 32) {
            above[k] = img[(int64_t)rr * L.Wimg + cr + min(k, wp - 1)];
            left[k] = img[(int64_t)(rr + min(k, hp - 1)) * L.Wimg + cr];
        }
        for (int p = lane; p < px; p += 32) target[p] = img[(int64_t)(r0 + (p >> shift)) * L.Wimg + c0 + (p & (W - 1))];
        __syncwarp();
        int dc = 0;
        for (int k = 1 + lane; k <= W; k += 32) dc += above[k] + left[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dc += __shfl_xor_sync(0xffffffffu, dc, o);
        dc = (dc + W) / (2 * W);                                        // :305-308
        int best_sse = 0x7fffffff, best_mode = 0;
        for (int mode = 0; mode < 35; ++mode) {
            const ModeInfo mi = mode >= 2 ? mode_info(mode) : ModeInfo{true, 0, 0};
            if (mode >= 2) {
                __syncwarp();
                build_ref_main(mi, W, above, left, ref_main, lane);
                __syncwarp();
            }
            int sse = 0;
            for (int p = lane; p < px; p += 32) {
                const int d = predict_pixel(mode, mi, p >> shift, p & (W - 1), W, shift, dc, above, left, ref_main) - (int)target[p];
                sse += d * d;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, o);
            if (sse < best_sse) {                                       // intraprediction.py:286: strictly better only
                best_sse = sse;
                best_mode = mode;
            }
        }
        if (lane == 0) {
            if (L.best_index) L.best_index[i] = (uint8_t)best_mode;
            if (L.psnr) L.psnr[i] = 10. * log10(255. * 255. / ((double)best_sse / (double)px + 1.e-6));   // tools/tools.py:364-401
        }
        if (L.pred) {
            const ModeInfo mi = best_mode >= 2 ? mode_info(best_mode) : ModeInfo{true, 0, 0};
            if (best_mode >= 2) {
                __syncwarp();
                build_ref_main(mi, W, above, left, ref_main, lane);
                __syncwarp();
            }
            for (int p = lane; p < px; p += 32) {
                L.pred[i * px + p] = (uint8_t)predict_pixel(best_mode, mi, p >> shift, p & (W - 1), W, shift, dc, above, left, ref_main);
            }
        }
        __syncwarp();
    }
}

}  // namespace

int launch_hevc_best_mode(const uint8_t* images, const int32_t* image_index, const int32_t* rows, const int32_t* cols, int64_t n,
                          int H, int Wimg, int W, int mask_w, int mask_h, uint8_t* best_index, double* psnr, uint8_t* pred,
                          cudaStream_t stream, int n_images) {
    if (n == 0) return 0;
    HevcLaunch L{images, image_index, rows, cols, n, H, Wimg, W, mask_w, mask_h, n_images, best_index, psnr, pred};
    int64_t grid = (n + WARPS - 1) / WARPS;
    if (grid > 148 * 16) grid = 148 * 16;
    hevc_best_mode_kernel<<<(unsigned)grid, WARPS * 32, 0, stream>>>(L);
    return 1;
}

}  // namespace pnn
