// HBM / latency-bound kernels of the PNN path: fused context gathers, the direct first convolution, channel-wise
// merger, col2im of the last transposed convolution with the fused epilogue, PSNR / win flags.
#include "kernels_common.cuh"

#include <cstdlib>

namespace pnn {

static inline int grid_for(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    const int64_t cap = 148 * 32;   // grid-stride beyond 32 waves of the 148 SMs
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// ---------------------------------------------------------------------------------------------
// Fused gather from uint8 images (reference sets/common.py:99-109 slicing, :454-461 mean + masks,
// :467-472 FC flattening = above row-major then left row-major).
// ---------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(256) gather_image_kernel(GatherLaunch L) {
    // thread = four consecutive pixels of one context row of one block: rows are 3W (above) or W (left) pixels long, both
    // multiples of four, and so are the masks, so a chunk is masked as a whole; one 8-byte store per bf16 plane (or one
    // 16-byte fp32 store) per thread instead of four 2-byte ones
    const int W = L.W;
    const int ca = 3 * W / 4, cl = W / 4;                    // chunks per row
    const int na4 = W * ca, per4 = na4 + 2 * W * cl;         // chunks per block: above, total
    const int64_t total = L.n * per4;
    for (int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x; gid < total; gid += (int64_t)gridDim.x * 256) {
        const int64_t i = gid / per4;
        const int q = (int)(gid - i * per4);
        const int img = L.image_index ? L.image_index[i] : 0;
        const int r0 = L.rows[i], c0 = L.cols[i];
        const bool img_ok = img >= 0 && img < L.n_images;      // (device-pointer entry point: the list is unchecked)
        const uint8_t* image = L.images + (int64_t)(img_ok ? img : 0) * L.H * L.Wimg;
        int r, c, e;
        bool masked;
        const bool is_above = q < na4;
        if (is_above) {
            const int rr = q / ca, cc = (q - rr * ca) * 4;
            r = r0 - W + rr;
            c = c0 - W + cc;
            masked = cc >= 3 * W - L.mask_w;
            e = rr * 3 * W + cc;
        } else {
            const int q2 = q - na4;
            const int rr = q2 / cl, cc = (q2 - rr * cl) * 4;
            r = r0 + rr;
            c = c0 - W + cc;
            masked = rr >= 2 * W - L.mask_h;
            e = rr * W + cc;
        }
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (img_ok && !masked && r >= 0 && r < L.H) {
            const uint8_t* src = image + (int64_t)r * L.Wimg;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c + j >= 0 && c + j < L.Wimg) v[j] = (float)src[c + j] - L.mean;
            }
        }
        const Act& dst = is_above ? L.above : L.left;
        const int64_t o = i * (is_above ? L.pitch_above : L.pitch_left) + e;
        if (SPLIT) {
            uint32_t hi[2], lo[2];
            split_bf16x2(v[0], v[1], hi[0], lo[0]);
            split_bf16x2(v[2], v[3], hi[1], lo[1]);
            *reinterpret_cast<uint2*>((__nv_bfloat16*)dst.p0 + o) = make_uint2(hi[0], hi[1]);
            *reinterpret_cast<uint2*>((__nv_bfloat16*)dst.p1 + o) = make_uint2(lo[0], lo[1]);
        } else {
            *reinterpret_cast<float4*>((float*)dst.p0 + o) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

int launch_gather_image(const GatherLaunch& L, cudaStream_t stream) {
    if (L.n == 0) return 0;
    const int grid = grid_for(L.n * (5 * L.W * L.W / 4), 256);
    if (L.split) gather_image_kernel<true><<<grid, 256, 0, stream>>>(L);
    else gather_image_kernel<false><<<grid, 256, 0, stream>>>(L);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// HM gather (reference extraction_context.cpp:56-205).  The host staged the raw int pixels of the
// L-shaped context; here: int -> float, minus mean, and the availability masking:
//   above columns [0, W)             always copied (Portion (1), extraction_context.cpp:119-127)
//   above columns W + i*unit_w ...   copied iff above unit i is available (:149-166)
//   left rows [0, left_rows_valid)   copied; the reference writer only advances on available units
//                                    (:189-205), so the copied rows are the first n_avail*unit_h ones.
// ---------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void gather_hm_kernel(GatherHmLaunch L) {
    const int W = L.W;
    const int na = 3 * W * W, total = 5 * W * W;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const float v = hm_context_value(L.staged, W, L.mean, e);
        if (e < na) act_store<SPLIT>(L.above, e, v);
        else act_store<SPLIT>(L.left, e - na, v);
    }
}

int launch_gather_hm(const GatherHmLaunch& L, cudaStream_t stream) {
    const int total = 5 * L.W * L.W;
    const int grid = (total + 255) / 256;
    if (L.split) gather_hm_kernel<true><<<grid, 256, 0, stream>>>(L);
    else gather_hm_kernel<false><<<grid, 256, 0, stream>>>(L);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// Weight tiles of the tcgen05 GEMM (layout in pnn_internal.h): one CTA per tile (nt, kb); element (row r = output column,
// kk = k within the 64-block) goes to byte r*128 + (((kk >> 3) ^ (r & 7)) << 4) + (kk & 7)*2 of the hi plane, the lo
// plane follows at + bn*128.  Rows / k beyond N / K are zero.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) make_tc_tiles_kernel(const float* __restrict__ w, int K, int N, uint8_t* __restrict__ tiles) {
    const int num_kb = (K + TC_BK - 1) / TC_BK;
    const int nt = blockIdx.x / num_kb, kb = blockIdx.x - nt * num_kb;
    int rem = N - nt * TC_BN;
    if (rem > TC_BN) rem = TC_BN;
    const int bn = (rem + 15) / 16 * 16;
    // every tile before the last N tile is TC_BN rows high
    uint8_t* hi = tiles + (size_t)nt * num_kb * 2 * TC_BN * 128 + (size_t)kb * 2 * bn * 128;
    uint8_t* lo = hi + (size_t)bn * 128;
    for (int idx = threadIdx.x; idx < bn * TC_BK; idx += 256) {
        const int kk = idx / bn, r = idx - kk * bn;          // r fastest: coalesced reads of w[k][n]
        const int n = nt * TC_BN + r, k = kb * TC_BK + kk;
        const float v = (n < N && k < K) ? w[(size_t)k * N + n] : 0.f;
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        const size_t o = (size_t)r * 128 + (size_t)(((kk >> 3) ^ (r & 7)) << 4) + (size_t)(kk & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(hi + o) = h;
        *reinterpret_cast<__nv_bfloat16*>(lo + o) = l;
    }
}

int launch_make_tc_tiles(const float* w_kn, int K, int N, uint8_t* tiles, cudaStream_t stream) {
    make_tc_tiles_kernel<<<tc_num_nt(N) * tc_num_kb(K), 256, 0, stream>>>(w_kn, K, N, tiles);
    return 1;
}

// ---------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void convert_input_kernel(const float* __restrict__ src, Act dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        act_store<SPLIT>(dst, i, src[i]);
    }
}

int launch_convert_input(const float* src, Act dst, int64_t n, int split, cudaStream_t stream) {
    if (n == 0) return 0;
    if (split) convert_input_kernel<true><<<grid_for(n, 256), 256, 0, stream>>>(src, dst, n);
    else convert_input_kernel<false><<<grid_for(n, 256), 256, 0, stream>>>(src, dst, n);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// First convolution of a branch, direct form (reference pnn/tfutils.py:134-139 with one input channel, then
// pnn/components.py:44-52 LeakyReLU).  Thread = PR output pixels of one column (rows PR*g .. PR*g + PR-1) x 16 output
// channels; lanes = consecutive columns.  The weights live in the kernel-parameter constant bank and every loop is
// unrolled, so the FFMAs take their weights through the uniform datapath (LDCU.128 + FFMA R, R, UR, R): no shared
// memory, no weight loads on the LSU, ((PR-1)*S + K) x K scalar input loads (L1) per PR*16*K*K FFMA.  One kernel per
// channel group keeps the straight-line code of a launch small (a switch over the group inside one kernel stalled on
// instruction fetch).  PR = 2 (96 registers, 20 warps / SM) measured faster than PR = 4 (166 registers, 12 warps).
// Fixed summation order (ky, kx ascending).  Output: bias, LeakyReLU, hi/lo split, one 32-byte store per pixel and plane.
// ---------------------------------------------------------------------------------------------
constexpr int CF_THREADS = 128;

__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

template <int K, int S, bool SPLIT, int CG, int PR>
__device__ __forceinline__ void conv_first_body(const ConvFirstLaunch& L, const ConvFirstWeights& Wt) {
    const unsigned OW = (unsigned)L.OW, groups = (unsigned)L.OH / PR;
    const int64_t total = (int64_t)L.n * groups * OW;
    constexpr int NR = (PR - 1) * S + K;                           // input rows feeding PR output rows
    for (int64_t t = (int64_t)blockIdx.x * CF_THREADS + threadIdx.x; t < total; t += (int64_t)gridDim.x * CF_THREADS) {
        const unsigned r0 = (unsigned)(t / OW), ox = (unsigned)(t - (int64_t)r0 * OW);
        const unsigned b = r0 / groups, g = r0 - b * groups;
        const float* inb = L.in + (int64_t)b * L.IH * L.IW;
        float acc[PR][16];
#pragma unroll
        for (int p = 0; p < PR; ++p)
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[p][c] = 0.f;
        const int ix0 = (int)ox * S - L.pad, iy0 = (int)g * PR * S - L.pad;
        // all NR x K input values of the item first (independent loads), then tap-major FFMAs: a weight is fetched once
        // and used for the PR pixels right away
        float v[NR][K];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int iy = iy0 + r;
            const bool row_ok = (unsigned)iy < (unsigned)L.IH;
            const float* row = inb + (row_ok ? iy : 0) * L.IW;
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const int ix = ix0 + kx;
                v[r][kx] = (row_ok && (unsigned)ix < (unsigned)L.IW) ? __ldg(row + ix) : 0.f;
            }
        }
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx)
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float w = Wt.w[(ky * K + kx) * 64 + CG * 16 + c];
#pragma unroll
                    for (int p = 0; p < PR; ++p) acc[p][c] = fmaf(v[p * S + ky][kx], w, acc[p][c]);
                }
#pragma unroll
        for (int p = 0; p < PR; ++p) {
            const int64_t pix = ((int64_t)b * L.OH + g * PR + p) * OW + ox;
            const int64_t o = pix * L.C + CG * 16;
            float y[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float z = acc[p][c] + Wt.b[CG * 16 + c];
                y[c] = fmaxf(z, 0.1f * z);
            }
            if (SPLIT) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    split_bf16x2(y[2 * j], y[2 * j + 1], hi[j], lo[j]);
                }
                st_global_v8((__nv_bfloat16*)L.out.p0 + o, hi);
                st_global_v8((__nv_bfloat16*)L.out.p1 + o, lo);
            } else {
                uint32_t u[8];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) u[j] = __float_as_uint(y[half * 8 + j]);
                    st_global_v8((float*)L.out.p0 + o + half * 8, u);
                }
            }
        }
    }
}

template <int K, int S, bool SPLIT, int CG, int PR>
__global__ void __launch_bounds__(CF_THREADS, PR == 4 ? 3 : 5) conv_first_kernel(const __grid_constant__ ConvFirstLaunch L,
                                                                                 const __grid_constant__ ConvFirstWeights Wt) {
    conv_first_body<K, S, SPLIT, CG, PR>(L, Wt);
}

// Batch-1 (in-loop) calls: all channel groups in ONE launch, group = blockIdx.y.  At this size the instruction-fetch
// stalls of the switch do not matter and three launches per branch are saved.
template <int K, int S, bool SPLIT, int PR>
__global__ void __launch_bounds__(CF_THREADS, PR == 4 ? 3 : 5) conv_first_all_groups_kernel(const __grid_constant__ ConvFirstLaunch L,
                                                                                            const __grid_constant__ ConvFirstWeights Wt) {
    switch (blockIdx.y) {
        case 0: conv_first_body<K, S, SPLIT, 0, PR>(L, Wt); break;
        case 1: conv_first_body<K, S, SPLIT, 1, PR>(L, Wt); break;
        case 2: conv_first_body<K, S, SPLIT, 2, PR>(L, Wt); break;
        default: conv_first_body<K, S, SPLIT, 3, PR>(L, Wt); break;
    }
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    // not volatile: the accumulator dependency alone fixes the summation order, independent chains may interleave
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// bf16x3 precision, batched calls: the same convolution as an im2col product [16 pixels x K*K] x [K*K x C] per warp on
// the tensor cores (mma.sync, register fragments: a tcgen05 operand would need the im2col rows written to shared memory
// in the K-major UMMA layout first, and with one input channel there is no 16-byte inner box for TMA).  ONE launch
// writes all C channels, so the input is read once instead of once per 16-channel group.  Per warp and tile of 16
// pixels: 8 (K = 3: 4) predicated LDG per lane and pixel row -- issued one tile AHEAD, so their latency hides behind
// the current tile --, hi/lo split of the taps, (C / 8) x ceil(K*K / 16) x 3 mma (hi*hi, hi*lo, lo*hi, k ascending:
// fixed order) on accumulators that start at the bias, LeakyReLU + split, a swizzled shared-memory transpose and
// 16-byte stores (a tile's 16 pixels are 2 KB contiguous in each plane).  The weight fragments are split once per CTA
// into shared memory in per-lane order (one conflict-free LDS.128 per (k step, channel octet)): in registers they cost
// 64 of them and a fourth of the resident warps.  Fragment layout: see merger_mma_kernel below.
struct FastDiv {                                   // n / d for n < 2^31 (host: fast_div_make)
    uint32_t mul, shift, d;
};
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv& f) { return f.d == 1 ? n : __umulhi(n, f.mul) >> f.shift; }
static FastDiv fast_div_make(uint32_t d) {
    FastDiv f{0, 0, d};
    if (d <= 1) return f;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;                   // ceil(log2 d) >= 1
    f.mul = (uint32_t)(((1ull << (31 + l)) / d) + 1);
    f.shift = l - 1;
    return f;
}

template <int K, int S, int NT>
__global__ void __launch_bounds__(CF_THREADS, 4) conv_first_mma_kernel(const __grid_constant__ ConvFirstLaunch L,
                                                                       const __grid_constant__ ConvFirstWeights Wt,
                                                                       const __grid_constant__ FastDiv divP,
                                                                       const __grid_constant__ FastDiv divOW) {
    constexpr int KK = K * K, KS = (KK + 15) / 16, NSLOT = KS * 4, WARPS = CF_THREADS / 32;
    __shared__ __align__(16) uint32_t stage[WARPS][2][16 * NT * 4];
    __shared__ __align__(16) uint4 wfrag[KS * NT * 32];            // {b0_hi, b1_hi, b0_lo, b1_lo} per (k step, octet, lane)
    __shared__ float2 sbias[NT * 4];                               // the accumulators start at the bias
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    for (int idx = threadIdx.x; idx < KS * NT * 32; idx += CF_THREADS) {
        const int ln = idx & 31, j = (idx >> 5) % NT, ks = idx / (32 * NT);
        const int n = j * 8 + (ln >> 2), k0 = ks * 16 + 2 * (ln & 3);
        float w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + (i & 1) + (i >> 1) * 8;
            w[i] = k < KK ? Wt.w[k * 64 + n] : 0.f;
        }
        uint4 f;
        split_bf16x2(w[0], w[1], f.x, f.z);
        split_bf16x2(w[2], w[3], f.y, f.w);
        wfrag[idx] = f;
    }
    // the K*K taps this lane feeds: slot 4*ks + {0, 1, 2, 3} = k 16*ks + 2t + {0, 1, 8, 9}; a tap is read when its row bit
    // (ky) and its column bit (8 + kx) are both set in the pixel's mask of in-range rows and columns
    int tap_off[NSLOT];
    uint32_t tap_need[NSLOT];
#pragma unroll
    for (int sl = 0; sl < NSLOT; ++sl) {
        const int k = (sl >> 2) * 16 + 2 * t + (sl & 1) + ((sl & 2) ? 8 : 0);
        const int ky = k / K, kx = k - ky * K;
        tap_need[sl] = k < KK ? (1u << ky | 1u << (8 + kx)) : 0x80000000u;    // a tap beyond K*K is never in range: value 0
        tap_off[sl] = ky * L.IW + kx;
    }
    if (threadIdx.x < NT * 4) sbias[threadIdx.x] = make_float2(Wt.b[2 * threadIdx.x], Wt.b[2 * threadIdx.x + 1]);
    __syncthreads();
    const uint32_t P = (uint32_t)(L.OH * L.OW), M = (uint32_t)L.n * P, tiles = (M + 15) >> 4;
    const uint32_t step = gridDim.x * WARPS;
    uint32_t* st_hi = stage[warp][0];
    uint32_t* st_lo = stage[warp][1];
    const uint4* wf = wfrag + lane;

    // the 2 x NSLOT taps of (tile, lane): rows g and g + 8
    auto load_taps = [&](uint32_t tile, float (&v)[2][NSLOT]) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t pix = (tile << 4) + g + 8 * r;
            const uint32_t b = fast_div(pix, divP), rem = pix - b * P;
            const uint32_t oy = fast_div(rem, divOW), ox = rem - oy * (uint32_t)L.OW;
            const int iy0 = (int)oy * S - L.pad, ix0 = (int)ox * S - L.pad;
            // rows max(0, -iy0) .. min(K, IH - iy0) - 1 and the same for the columns are inside the input
            const uint32_t rows = (0xffffffffu << max(0, -iy0)) & ~(0xffffffffu << min(K, L.IH - iy0));
            const uint32_t cols = (0xffffffffu << max(0, -ix0)) & ~(0xffffffffu << min(K, L.IW - ix0));
            const uint32_t have = pix < M ? (rows | cols << 8) : 0u;
            const int off = (int)b * (L.IH * L.IW) + iy0 * L.IW + ix0;     // launcher: n * IH * IW < 2^31
#pragma unroll
            for (int sl = 0; sl < NSLOT; ++sl) v[r][sl] = (~have & tap_need[sl]) == 0u ? __ldg(L.in + (off + tap_off[sl])) : 0.f;
        }
    };

    uint32_t tile = blockIdx.x * WARPS + warp;
    float v[2][NSLOT];
    if (tile < tiles) load_taps(tile, v);
    for (; tile < tiles; tile += step) {
        uint32_t ah[KS][4], al[KS][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                split_bf16x2(v[r][4 * ks], v[r][4 * ks + 1], ah[ks][r], al[ks][r]);
                split_bf16x2(v[r][4 * ks + 2], v[r][4 * ks + 3], ah[ks][2 + r], al[ks][2 + r]);
            }
        if (tile + step < tiles) load_taps(tile + step, v);        // in flight during this tile's products and stores
        __syncwarp();                                              // the previous tile's staged rows have been read
#pragma unroll
        for (int j0 = 0; j0 < NT; j0 += 4) {                       // 4 independent accumulator chains in flight
            float acc[4][4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float2 bj = sbias[(j0 + jj) * 4 + t];
                acc[jj][0] = bj.x; acc[jj][1] = bj.y; acc[jj][2] = bj.x; acc[jj][3] = bj.y;
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                uint4 b[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) b[jj] = wf[(ks * NT + j0 + jj) * 32];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) mma_bf16_16816(acc[jj], ah[ks], b[jj].x, b[jj].y);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) mma_bf16_16816(acc[jj], ah[ks], b[jj].z, b[jj].w);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) mma_bf16_16816(acc[jj], al[ks], b[jj].x, b[jj].y);
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    uint32_t hi, lo;
                    split_bf16x2(fmaxf(acc[jj][2 * r], 0.1f * acc[jj][2 * r]), fmaxf(acc[jj][2 * r + 1], 0.1f * acc[jj][2 * r + 1]), hi, lo);
                    const int row = g + 8 * r, word = ((j0 + jj) * 4 + t) ^ ((g & (NT - 1)) << 2);   // 16-byte chunk index XOR row: no bank conflicts at C = 64
                    st_hi[row * (NT * 4) + word] = hi;
                    st_lo[row * (NT * 4) + word] = lo;
                }
        }
        __syncwarp();
        constexpr int CHUNKS = NT;                                 // 16-byte chunks per row and plane
#pragma unroll
        for (int it = 0; it < 16 * CHUNKS / 32; ++it) {
            const int idx = it * 32 + lane, row = idx / CHUNKS, chunk = idx - row * CHUNKS;
            const uint32_t pix = (tile << 4) + row;
            if (pix < M) {
                const int phys = chunk ^ (row & (NT - 1));
                const uint4 vh = *reinterpret_cast<const uint4*>(st_hi + row * (NT * 4) + phys * 4);
                const uint4 vl = *reinterpret_cast<const uint4*>(st_lo + row * (NT * 4) + phys * 4);
                *reinterpret_cast<uint4*>((__nv_bfloat16*)L.out.p0 + (int64_t)pix * (NT * 8) + chunk * 8) = vh;
                *reinterpret_cast<uint4*>((__nv_bfloat16*)L.out.p1 + (int64_t)pix * (NT * 8) + chunk * 8) = vl;
            }
        }
    }
}

template <int K, int S, bool SPLIT, int PR>
void conv_first_launch_groups(const ConvFirstLaunch& L, const ConvFirstWeights& W, unsigned grid, cudaStream_t stream) {
    if (grid <= 148) {
        conv_first_all_groups_kernel<K, S, SPLIT, PR><<<dim3(grid, L.C / 16), CF_THREADS, 0, stream>>>(L, W);
        return;
    }
    conv_first_kernel<K, S, SPLIT, 0, PR><<<grid, CF_THREADS, 0, stream>>>(L, W);
    conv_first_kernel<K, S, SPLIT, 1, PR><<<grid, CF_THREADS, 0, stream>>>(L, W);
    if (L.C == 64) {
        conv_first_kernel<K, S, SPLIT, 2, PR><<<grid, CF_THREADS, 0, stream>>>(L, W);
        conv_first_kernel<K, S, SPLIT, 3, PR><<<grid, CF_THREADS, 0, stream>>>(L, W);
    }
}

template <int K, int S>
static void conv_first_launch_mma(const ConvFirstLaunch& L, const ConvFirstWeights& W, cudaStream_t stream) {
    const int64_t tiles = ((int64_t)L.n * L.OH * L.OW + 15) >> 4;
    const unsigned grid = (unsigned)std::min<int64_t>((tiles + CF_THREADS / 32 - 1) / (CF_THREADS / 32), 148 * 4);   // persistent: 4 CTAs / SM
    const FastDiv dp = fast_div_make((uint32_t)(L.OH * L.OW)), dw = fast_div_make((uint32_t)L.OW);
    if (L.C == 64) conv_first_mma_kernel<K, S, 8><<<grid, CF_THREADS, 0, stream>>>(L, W, dp, dw);
    else conv_first_mma_kernel<K, S, 4><<<grid, CF_THREADS, 0, stream>>>(L, W, dp, dw);
}

int launch_conv_first(const ConvFirstLaunch& L, const ConvFirstWeights& W, cudaStream_t stream) {
    if (L.n == 0) return 0;
    if (L.split && !L.in_loop && (int64_t)L.n * L.OH * L.OW < (1ll << 31) - 16 && (int64_t)L.n * L.IH * L.IW < (1ll << 31)) {   // every batched bf16x3 call, whatever its size: one arithmetic per block
        if (L.k == 5 && L.stride == 2) conv_first_launch_mma<5, 2>(L, W, stream);
        else conv_first_launch_mma<3, 1>(L, W, stream);
        return 1;
    }
    constexpr int PR = 2;
    const int64_t total = (int64_t)L.n * (L.OH / PR) * L.OW;
    const unsigned grid = (unsigned)std::min<int64_t>((total + CF_THREADS - 1) / CF_THREADS, 148 * 20);
    if (L.k == 5 && L.stride == 2) {
        if (L.split) conv_first_launch_groups<5, 2, true, PR>(L, W, grid, stream);
        else conv_first_launch_groups<5, 2, false, PR>(L, W, grid, stream);
    } else {
        if (L.split) conv_first_launch_groups<3, 1, true, PR>(L, W, grid, stream);
        else conv_first_launch_groups<3, 1, false, PR>(L, W, grid, stream);
    }
    return grid <= 148 ? 1 : L.C / 16;
}

// ---------------------------------------------------------------------------------------------
// col2im of the last transposed convolution, fused with the output epilogue (reference
// pnn/tfutils.py:455-462; TComPrediction.cpp:621-635 / tools/tools.py:49).  Gather form, fixed order
// (ky, kx ascending): out[y, x] = bias + sum_{ky, kx : (y+pad-ky) % s == 0, ...} D[((y+pad-ky)/s, (x+pad-kx)/s), ky*k+kx].
// ---------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(128) col2im_kernel(Col2imLaunch L) {
    const int OH = L.IH * L.stride, OW = L.IW * L.stride, P = OH * OW;
    const int tiles = (P + 127) >> 7;
    const int b = blockIdx.x / tiles, tile = blockIdx.x - b * tiles;
    const int pix = (tile << 7) + threadIdx.x;
    if (pix >= P) return;
    const int y = pix / OW, x = pix - y * OW;
    const int64_t base = (int64_t)b * L.IH * L.IW;
    float acc = 0.f;
    for (int ky = 0; ky < L.k; ++ky) {
        const int ty = y + L.pad - ky;
        if (ty < 0 || (L.stride == 2 && (ty & 1))) continue;
        const int iy = L.stride == 2 ? ty >> 1 : ty;
        if (iy >= L.IH) continue;
        for (int kx = 0; kx < L.k; ++kx) {
            const int tx = x + L.pad - kx;
            if (tx < 0 || (L.stride == 2 && (tx & 1))) continue;
            const int ix = L.stride == 2 ? tx >> 1 : tx;
            if (ix >= L.IW) continue;
            acc += act_load<SPLIT>(L.d, (base + iy * L.IW + ix) * L.NP + ky * L.k + kx);
        }
    }
    final_store(L.fin, (int64_t)b * P + pix, acc + L.bias);
}

// The same with the sample's per-tap products staged in shared memory first (maps up to 16 x 16 input pixels, i.e. every net
// but CONV-64): the global reads become contiguous 8- / 16-byte loads instead of one scattered 2-byte load per tap and plane
// (1.0 -> ms per bench step).  Same tap order, same arithmetic.
constexpr int C2I_MAX = 8192;

template <bool SPLIT>
__global__ void __launch_bounds__(256) col2im_smem_kernel(Col2imLaunch L) {
    __shared__ __align__(16) float d_s[C2I_MAX];
    const int n_in = L.IH * L.IW * L.NP;                      // a multiple of 16
    const int64_t base = (int64_t)blockIdx.x * n_in;
    for (int idx = threadIdx.x * 4; idx < n_in; idx += 256 * 4) {
        float4 v;
        if (SPLIT) {
            const uint2 h = __ldg(reinterpret_cast<const uint2*>((const __nv_bfloat16*)L.d.p0 + base + idx));
            const uint2 l = __ldg(reinterpret_cast<const uint2*>((const __nv_bfloat16*)L.d.p1 + base + idx));
            v.x = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
            v.y = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
            v.z = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
            v.w = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
        } else {
            v = __ldg(reinterpret_cast<const float4*>((const float*)L.d.p0 + base + idx));
        }
        *reinterpret_cast<float4*>(d_s + idx) = v;
    }
    __syncthreads();
    const int OH = L.IH * L.stride, OW = L.IW * L.stride, P = OH * OW;
    for (int pix = threadIdx.x; pix < P; pix += 256) {
        const int y = pix / OW, x = pix - y * OW;
        float acc = 0.f;
        for (int ky = 0; ky < L.k; ++ky) {
            const int ty = y + L.pad - ky;
            if (ty < 0 || (L.stride == 2 && (ty & 1))) continue;
            const int iy = L.stride == 2 ? ty >> 1 : ty;
            if (iy >= L.IH) continue;
            for (int kx = 0; kx < L.k; ++kx) {
                const int tx = x + L.pad - kx;
                if (tx < 0 || (L.stride == 2 && (tx & 1))) continue;
                const int ix = L.stride == 2 ? tx >> 1 : tx;
                if (ix >= L.IW) continue;
                acc += d_s[(iy * L.IW + ix) * L.NP + ky * L.k + kx];
            }
        }
        final_store(L.fin, (int64_t)blockIdx.x * P + pix, acc + L.bias);
    }
}

int launch_col2im(const Col2imLaunch& L, cudaStream_t stream) {
    if (L.n == 0) return 0;
    if (L.IH * L.IW * L.NP <= C2I_MAX) {
        if (L.split) col2im_smem_kernel<true><<<(unsigned)L.n, 256, 0, stream>>>(L);
        else col2im_smem_kernel<false><<<(unsigned)L.n, 256, 0, stream>>>(L);
        return 1;
    }
    const int P = L.IH * L.stride * L.IW * L.stride;
    const int64_t grid = (int64_t)L.n * ((P + 127) / 128);
    if (L.split) col2im_kernel<true><<<(unsigned)grid, 128, 0, stream>>>(L);
    else col2im_kernel<false><<<(unsigned)grid, 128, 0, stream>>>(L);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// Channel-wise fully-connected merger + LeakyReLU (reference pnn/tfutils.py:60-73,
// pnn/components.py:225-231): per channel c, the 48 values of the above map (row-major) followed by
// the 32 values of the left map are fully connected to 16 outputs.  One thread per (sample, channel),
// channel fastest; the weights were transposed on the host to [80][16][C] so that loads coalesce.
// ---------------------------------------------------------------------------------------------
// v4: persistent CTAs, each owns a group of 16 channels whose 80x16 weights stay in shared memory
// ([q][channel][16 outputs + 4 pad floats]) and loops over tiles of 128 samples.  Thread = (8 channels, one
// sample): the 8 channel values of an input pixel arrive with one 16-byte load per plane (several q in flight,
// the kernel is otherwise bound by load latency), the weights with warp-broadcast LDS.128 (a warp works on a
// single channel octet), and the 8 x 16 accumulators live in registers.  Fixed order: q ascending.
constexpr int MG_CH = 16, MG_ROW = 20, MG_TILE = 128, MG_UNROLL = 4;
constexpr int MG_SMEM = (80 * MG_CH * MG_ROW + 16 * MG_CH) * (int)sizeof(float);

template <bool SPLIT>
__device__ __forceinline__ void merger_load8(const Act& a, int64_t idx, uint4& h, uint4& l) {
    if (SPLIT) {
        h = __ldg((const uint4*)((const __nv_bfloat16*)a.p0 + idx));
        l = __ldg((const uint4*)((const __nv_bfloat16*)a.p1 + idx));
    } else {
        h = __ldg((const uint4*)((const float*)a.p0 + idx));
        l = __ldg((const uint4*)((const float*)a.p0 + idx + 4));
    }
}
template <bool SPLIT>
__device__ __forceinline__ void merger_unpack8(const uint4& h, const uint4& l, float (&x)[8]) {
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
    if (SPLIT) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[2 * j] = __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
            x[2 * j + 1] = __uint_as_float(hw[j] & 0xffff0000u) + __uint_as_float(lw[j] & 0xffff0000u);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[j] = __uint_as_float(hw[j]);
            x[4 + j] = __uint_as_float(lw[j]);
        }
    }
}

template <bool SPLIT>
__global__ void __launch_bounds__(256, 1) merger_kernel(MergerLaunch L, int groups_per_cg) {
    extern __shared__ float mg_s[];
    float* w_s = mg_s;                                  // [80][16 ch][20]
    float* b_s = mg_s + 80 * MG_CH * MG_ROW;            // [16 ch][16]
    const int ncg = L.C / MG_CH;
    const int cg = blockIdx.x % ncg, grp = blockIdx.x / ncg;
    for (int i = threadIdx.x; i < 80 * 16 * MG_CH; i += 256) {
        const int ch = i % MG_CH, qp = i / MG_CH;       // L.w is [80][16][C]: coalesced over the channel
        const int q = qp >> 4, p = qp & 15;
        w_s[(q * MG_CH + ch) * MG_ROW + p] = L.w[(int64_t)qp * L.C + cg * MG_CH + ch];
    }
    for (int i = threadIdx.x; i < 16 * MG_CH; i += 256) {
        const int ch = i % MG_CH, p = i / MG_CH;
        b_s[ch * 16 + p] = L.bias[p * L.C + cg * MG_CH + ch];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int oct = warp & 1;                           // channel octet of the group: warp-uniform
    const int c0 = cg * MG_CH + oct * 8;
    const int tiles = (L.n + MG_TILE - 1) / MG_TILE;
    for (int tile = grp; tile < tiles; tile += groups_per_cg) {
        const int64_t b = (int64_t)tile * MG_TILE + (warp >> 1) * 32 + lane;
        const bool ok = b < L.n;
        const int64_t bb = ok ? b : 0;
        float acc[8][16];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
#pragma unroll
            for (int p = 0; p < 16; ++p) acc[ch][p] = 0.f;
        for (int q0 = 0; q0 < 80; q0 += MG_UNROLL) {     // 48 and 80 are multiples of MG_UNROLL
            uint4 h[MG_UNROLL], l[MG_UNROLL];
#pragma unroll
            for (int u = 0; u < MG_UNROLL; ++u) {
                const int q = q0 + u;
                if (q < 48) merger_load8<SPLIT>(L.in0, (bb * 48 + q) * L.C + c0, h[u], l[u]);
                else merger_load8<SPLIT>(L.in1, (bb * 32 + (q - 48)) * L.C + c0, h[u], l[u]);
            }
#pragma unroll
            for (int u = 0; u < MG_UNROLL; ++u) {
                float x[8];
                merger_unpack8<SPLIT>(h[u], l[u], x);
                const float* wq = w_s + ((q0 + u) * MG_CH + oct * 8) * MG_ROW;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const float4* wr = (const float4*)(wq + ch * MG_ROW);
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float4 w4 = wr[v];
                        acc[ch][4 * v + 0] = fmaf(x[ch], w4.x, acc[ch][4 * v + 0]);
                        acc[ch][4 * v + 1] = fmaf(x[ch], w4.y, acc[ch][4 * v + 1]);
                        acc[ch][4 * v + 2] = fmaf(x[ch], w4.z, acc[ch][4 * v + 2]);
                        acc[ch][4 * v + 3] = fmaf(x[ch], w4.w, acc[ch][4 * v + 3]);
                    }
                }
            }
        }
        if (!ok) continue;
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            float y[8];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) y[ch] = leaky_relu(acc[ch][p] + b_s[(oct * 8 + ch) * 16 + p]);
            const int64_t o = (b * 16 + p) * L.C + c0;
            if (SPLIT) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    split_bf16x2(y[2 * j], y[2 * j + 1], hi[j], lo[j]);
                }
                *(uint4*)((__nv_bfloat16*)L.out.p0 + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *(uint4*)((__nv_bfloat16*)L.out.p1 + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            } else {
                float4* po = (float4*)((float*)L.out.p0 + o);
                po[0] = make_float4(y[0], y[1], y[2], y[3]);
                po[1] = make_float4(y[4], y[5], y[6], y[7]);
            }
        }
    }
}

static int g_merger_init = 0;
constexpr int MG5_THREADS = 128;
constexpr int MG5_SMEM = 2 * 5 * 8 * 2 * 32 * 16;    // [octet][k16 chunk][channel][n tile][lane] x {b0_hi, b1_hi, b0_lo, b1_lo}
__global__ void merger_mma_kernel(MergerLaunch L, int groups);

void small_kernels_init() {
    cudaFuncSetAttribute(merger_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG5_SMEM);
    cudaFuncSetAttribute(merger_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM);
    cudaFuncSetAttribute(merger_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM);
    g_merger_init = 1;
}

// v5 (bf16x3 precision): the same per-channel [samples x 80] x [80 x 16] products on the tensor cores, through the
// warp-level mma.sync path because its operands are REGISTER fragments: the activations are channel-last ([sample, q, C]),
// so a K-major shared-memory operand for one channel (what tcgen05 needs) would cost a transpose through shared memory,
// whereas here a lane's 16-byte load (8 channels of one (sample, q)) feeds 8 channel problems after one PRMT per
// fragment register.  The op is bound by HBM (98 KB per sample at C = 256 against 0.65 MFLOP), the FFMA version (v4,
// kept for the fp32 yard-stick) was bound by the LSU: 32 broadcast LDS.128 per 128 FFMA, ncu l1tex 88 %, FMA pipe 23 %.
// Warp = 16 samples x 8 channels: per 16 positions 16 LDG.128 + 16 LDS.128 (weight fragments, prepared once per CTA in
// the exact per-lane order) + 64 PRMT + 48 mma.m16n8k16 (hi*hi, hi*lo, lo*hi; fp32 accumulate, q ascending).
// CTA = 4 warps = 2 channel octets x 2 sample halves, so that both halves of every 32-byte sector are consumed by the
// same SM; persistent over sample tiles.  Fragment layout (PTX ISA, m16n8k16 .bf16): g = lane >> 2, t = lane & 3;
// A: a0 (row g, k 2t..2t+1), a1 (row g+8, same k), a2 (row g, k 2t+8..2t+9), a3 (row g+8, k 2t+8..); B: b0 (k 2t..2t+1,
// n g), b1 (k 2t+8.., n g); C: c0, c1 (row g, n 2t, 2t+1), c2, c3 (row g+8, same n).
__device__ __forceinline__ uint32_t u4_word(const uint4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

__global__ void __launch_bounds__(MG5_THREADS, 2) merger_mma_kernel(MergerLaunch L, int groups) {
    extern __shared__ __align__(16) uint8_t mg5_smem[];
    uint4* wfrag = reinterpret_cast<uint4*>(mg5_smem);
    const int C = L.C;
    const int pair = blockIdx.x / groups, grp = blockIdx.x - pair * groups;    // channel group of 16, sample-tile group
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    // weight fragments of the CTA's 16 channels: hi/lo split once
    for (int idx = threadIdx.x; idx < 2 * 5 * 8 * 2 * 32; idx += MG5_THREADS) {
        const int ln = idx & 31, nt = (idx >> 5) & 1, j = (idx >> 6) & 7, kc = (idx >> 9) % 5, oct = idx / (5 * 512);
        const int gg = ln >> 2, tt = ln & 3;
        const int c = pair * 16 + oct * 8 + j, n = nt * 8 + gg;
        const int ks[4] = {kc * 16 + 2 * tt, kc * 16 + 2 * tt + 1, kc * 16 + 2 * tt + 8, kc * 16 + 2 * tt + 9};
        float w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = __ldg(L.w + ((int64_t)ks[i] * 16 + n) * C + c);
        uint4 f;
        split_bf16x2(w[0], w[1], f.x, f.z);
        split_bf16x2(w[2], w[3], f.y, f.w);
        wfrag[idx] = f;
    }
    __syncthreads();
    const int oct = warp & 1, half = warp >> 1;
    const int c0 = pair * 16 + oct * 8;
    const uint4* wf = wfrag + oct * (5 * 512) + lane;
    const __nv_bfloat16* in0_hi = (const __nv_bfloat16*)L.in0.p0 + c0;
    const __nv_bfloat16* in0_lo = (const __nv_bfloat16*)L.in0.p1 + c0;
    const __nv_bfloat16* in1_hi = (const __nv_bfloat16*)L.in1.p0 + c0;
    const __nv_bfloat16* in1_lo = (const __nv_bfloat16*)L.in1.p1 + c0;
    const int tiles = (L.n + 31) >> 5;                        // 32 samples per CTA iteration
    for (int tile = grp; tile < tiles; tile += groups) {
        const int s_base = tile * 32 + half * 16;
        if (s_base >= L.n) continue;                          // warp-uniform
        // the two samples (fragment rows g and g + 8) of this lane, clamped for the loads
        const int sa = min(s_base + g, L.n - 1), sb = min(s_base + g + 8, L.n - 1);
        float acc[8][2][4];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][nt][i] = 0.f;
#pragma unroll
        for (int kc = 0; kc < 5; ++kc) {
            // positions q = kc*16 + {2t, 2t+1, 2t+8, 2t+9}: q < 48 lies in the above map (48 positions), else in the left map
            const bool above = kc < 3;
            const __nv_bfloat16* ph = above ? in0_hi : in1_hi;
            const __nv_bfloat16* pl = above ? in0_lo : in1_lo;
            const int npos = above ? 48 : 32, q0 = (above ? kc * 16 : (kc - 3) * 16) + 2 * t;
            const int64_t ra = ((int64_t)sa * npos + q0) * C, rb = ((int64_t)sb * npos + q0) * C;
            // x[row][i]: row 0 = sample g, row 1 = sample g + 8; i = position 2t, 2t+1, 2t+8, 2t+9
            uint4 xh[2][4], xl[2][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t dq = (int64_t)((i & 1) + (i >> 1) * 8) * C;
                xh[0][i] = __ldg((const uint4*)(ph + ra + dq));
                xh[1][i] = __ldg((const uint4*)(ph + rb + dq));
                xl[0][i] = __ldg((const uint4*)(pl + ra + dq));
                xl[1][i] = __ldg((const uint4*)(pl + rb + dq));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
                uint32_t ah[4], al[4];
                ah[0] = __byte_perm(u4_word(xh[0][0], j >> 1), u4_word(xh[0][1], j >> 1), sel);
                ah[1] = __byte_perm(u4_word(xh[1][0], j >> 1), u4_word(xh[1][1], j >> 1), sel);
                ah[2] = __byte_perm(u4_word(xh[0][2], j >> 1), u4_word(xh[0][3], j >> 1), sel);
                ah[3] = __byte_perm(u4_word(xh[1][2], j >> 1), u4_word(xh[1][3], j >> 1), sel);
                al[0] = __byte_perm(u4_word(xl[0][0], j >> 1), u4_word(xl[0][1], j >> 1), sel);
                al[1] = __byte_perm(u4_word(xl[1][0], j >> 1), u4_word(xl[1][1], j >> 1), sel);
                al[2] = __byte_perm(u4_word(xl[0][2], j >> 1), u4_word(xl[0][3], j >> 1), sel);
                al[3] = __byte_perm(u4_word(xl[1][2], j >> 1), u4_word(xl[1][3], j >> 1), sel);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const uint4 b = wf[((kc * 8 + j) * 2 + nt) * 32];      // {b0_hi, b1_hi, b0_lo, b1_lo}
                    mma_bf16_16816(acc[j][nt], ah, b.x, b.y);               // hi * hi
                    mma_bf16_16816(acc[j][nt], ah, b.z, b.w);               // hi * lo
                    mma_bf16_16816(acc[j][nt], al, b.x, b.y);               // lo * hi
                }
            }
        }
        // epilogue: (sample, output) pairs of this lane: rows {g, g+8} x outputs {2t, 2t+1, 8+2t, 8+2t+1}; the 8 channels of
        // a pair are the 8 accumulator sets -> one 16-byte store per plane
#pragma unroll
        for (int row = 0; row < 2; ++row) {
            const int smp = s_base + g + 8 * row;
            if (smp >= L.n) continue;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int o = nt * 8 + 2 * t + e;
                    const float4 b0 = __ldg((const float4*)(L.bias + (int64_t)o * C + c0));
                    const float4 b1 = __ldg((const float4*)(L.bias + (int64_t)o * C + c0 + 4));
                    const float bias[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    float y[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = leaky_relu(acc[j][nt][2 * row + e] + bias[j]);
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) split_bf16x2(y[2 * j], y[2 * j + 1], hi[j], lo[j]);
                    const int64_t off = ((int64_t)smp * 16 + o) * C + c0;
                    *(uint4*)((__nv_bfloat16*)L.out.p0 + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *(uint4*)((__nv_bfloat16*)L.out.p1 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
    }
}

// In-loop (batch-1) merger: thread = (sample, output o, channel c), c fastest, 80-term fp32 dot product in fixed order.
// For one sample the tensor-core kernel spends its 21 us preparing 80 KB of weight fragments per CTA; this one reads the
// 80*16*C weights once, coalesced.
template <bool SPLIT>
__global__ void __launch_bounds__(256) merger_small_kernel(MergerLaunch L) {
    const int C = L.C;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= L.n * 16 * C) return;
    const int c = idx % C, o = (idx / C) & 15, s = idx / (16 * C);
    float acc = 0.f;
    const int64_t b0 = (int64_t)s * 48 * C + c, b1 = (int64_t)s * 32 * C + c;
#pragma unroll 8
    for (int q = 0; q < 48; ++q) acc = fmaf(act_load<SPLIT>(L.in0, b0 + (int64_t)q * C), __ldg(L.w + ((int64_t)q * 16 + o) * C + c), acc);
#pragma unroll 8
    for (int q = 0; q < 32; ++q) acc = fmaf(act_load<SPLIT>(L.in1, b1 + (int64_t)q * C), __ldg(L.w + ((int64_t)(48 + q) * 16 + o) * C + c), acc);
    act_store<SPLIT>(L.out, ((int64_t)s * 16 + o) * C + c, leaky_relu(acc + __ldg(L.bias + (int64_t)o * C + c)));
}

int launch_merger(const MergerLaunch& L, cudaStream_t stream) {
    if (L.n == 0) return 0;
    if (L.C % MG_CH) return -1;
    if (!g_merger_init) {
        cudaFuncSetAttribute(merger_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG5_SMEM);
        cudaFuncSetAttribute(merger_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM);
        cudaFuncSetAttribute(merger_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM);
        g_merger_init = 1;
    }
    if (L.in_loop) {
        if (L.split) merger_small_kernel<true><<<(L.n * 16 * L.C + 255) / 256, 256, 0, stream>>>(L);
        else merger_small_kernel<false><<<(L.n * 16 * L.C + 255) / 256, 256, 0, stream>>>(L);
        return 1;
    }
    static const int use_mma = getenv("PNN_MERGER_MMA") ? atoi(getenv("PNN_MERGER_MMA")) != 0 : 1;
    if (L.split && use_mma) {
        const int pairs = L.C / 16, tiles32 = (L.n + 31) / 32;
        int groups5 = (2 * 148 + pairs - 1) / pairs;        // about two resident CTAs per SM
        if (groups5 > tiles32) groups5 = tiles32;
        merger_mma_kernel<<<pairs * groups5, MG5_THREADS, MG5_SMEM, stream>>>(L, groups5);
        return 1;
    }
    const int ncg = L.C / MG_CH;
    const int tiles = (L.n + MG_TILE - 1) / MG_TILE;
    int groups = (148 + ncg - 1) / ncg;                 // about one resident CTA per SM
    if (groups > tiles) groups = tiles;
    if (groups < 1) groups = 1;
    if (L.split) merger_kernel<true><<<ncg * groups, 256, MG_SMEM, stream>>>(L, groups);
    else merger_kernel<false><<<ncg * groups, 256, MG_SMEM, stream>>>(L, groups);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// PSNR per block (reference tools/tools.py:364-401): 10*log10(255^2 / (mean((a-b)^2) + 1e-6)), float64.
// One warp per block; the sum of squared differences is an exact integer.
// ---------------------------------------------------------------------------------------------
__global__ void psnr_kernel(const uint8_t* __restrict__ images, const int32_t* __restrict__ image_index,
                            const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int64_t n,
                            int H, int Wimg, int W, const uint8_t* __restrict__ pred, double* __restrict__ out, int n_images) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n; i += nwarps) {
        const int img = image_index ? image_index[i] : 0;
        // a target block that leaves the image (only possible through the unchecked device-pointer entry point): PSNR NaN
        if (img < 0 || img >= n_images || rows[i] < 0 || cols[i] < 0 || rows[i] + W > H || cols[i] + W > Wimg) {
            if (lane == 0) out[i] = __longlong_as_double(0x7ff8000000000000LL);
            continue;
        }
        const uint8_t* base = images + ((int64_t)img * H + rows[i]) * Wimg + cols[i];
        const uint8_t* p = pred + i * W * W;
        int sse = 0;
        for (int e = lane; e < W * W; e += 32) {
            const int r = e / W, c = e - r * W;
            const int d = (int)base[(int64_t)r * Wimg + c] - (int)p[e];
            sse += d * d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, o);
        if (lane == 0) {
            const double mse = (double)sse / (double)(W * W);
            out[i] = 10. * log10(255. * 255. / (mse + 1.e-6));
        }
    }
}

// Blocks of 4 x 4 and 8 x 8 pixels: one THREAD per block (a warp per block leaves half / none of its lanes busy and the
// bench has 1.2 M such blocks).  Same integer sum, same formula.
__global__ void __launch_bounds__(256) psnr_small_kernel(const uint8_t* __restrict__ images, const int32_t* __restrict__ image_index,
                                                         const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int64_t n,
                                                         int H, int Wimg, int W, const uint8_t* __restrict__ pred, double* __restrict__ out,
                                                         int n_images) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int img = image_index ? image_index[i] : 0;
        if (img < 0 || img >= n_images || rows[i] < 0 || cols[i] < 0 || rows[i] + W > H || cols[i] + W > Wimg) {
            out[i] = __longlong_as_double(0x7ff8000000000000LL);
            continue;
        }
        const uint8_t* base = images + ((int64_t)img * H + rows[i]) * Wimg + cols[i];
        const uint32_t* p = reinterpret_cast<const uint32_t*>(pred + i * W * W);     // W*W is a multiple of 16 bytes
        int sse = 0;
        for (int r = 0; r < W; ++r) {
            const uint8_t* row = base + (int64_t)r * Wimg;
            for (int c4 = 0; c4 < W; c4 += 4) {
                const uint32_t q = __ldg(p + (r * W + c4) / 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int d = (int)row[c4 + j] - (int)((q >> (8 * j)) & 0xffu);
                    sse += d * d;
                }
            }
        }
        const double mse = (double)sse / (double)(W * W);
        out[i] = 10. * log10(255. * 255. / (mse + 1.e-6));
    }
}

int launch_psnr(const uint8_t* images, const int32_t* image_index, const int32_t* rows, const int32_t* cols,
                int64_t n, int H, int Wimg, int W, const uint8_t* pred_u8, double* out, cudaStream_t stream, int n_images) {
    if (n == 0) return 0;
    if (W <= 8) {
        psnr_small_kernel<<<grid_for(n, 256), 256, 0, stream>>>(images, image_index, rows, cols, n, H, Wimg, W, pred_u8, out, n_images);
        return 1;
    }
    psnr_kernel<<<grid_for(n * 32, 256), 256, 0, stream>>>(images, image_index, rows, cols, n, H, Wimg, W, pred_u8, out, n_images);
    return 1;
}

// reference comparing_pnn_ipfcns_hevc_best_mode.py:87: numpy.count_nonzero(psnrs_nn - psnrs_hevc_best_mode > 0.)
__global__ void win_flags_kernel(const double* __restrict__ psnr, const double* __restrict__ base, int64_t n,
                                 uint8_t* __restrict__ win) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        win[i] = (psnr[i] - base[i] > 0.) ? 1 : 0;
    }
}

int launch_win_flags(const double* psnr, const double* baseline, int64_t n, uint8_t* win, cudaStream_t stream) {
    if (n == 0) return 0;
    win_flags_kernel<<<grid_for(n, 256), 256, 0, stream>>>(psnr, baseline, n, win);
    return 1;
}

}  // namespace pnn
