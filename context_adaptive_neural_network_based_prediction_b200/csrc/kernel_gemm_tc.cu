// tcgen05 implicit-GEMM kernel for sm_100a: the throughput path of every GEMM-shaped PNN layer
// (FC layers, convolutions, transposed-convolution phases).
//
// Arithmetic: "bf16x3".  Activations and weights are kept as two bf16 planes (hi, lo) with
// value = hi + lo (16 mantissa bits).  Each 16-wide K step issues three tcgen05.mma.kind::f16:
//     D += A_hi * B_hi,  D += A_hi * B_lo,  D += A_lo * B_hi        (fp32 accumulate in TMEM)
// which keeps the max error ~5e-4 pixel units at outputs of ~50 (1e-2 is the parity bound;
// a single bf16 or tf32 pass misses it, see DESIGN.md "Precision").  The reduction order is fixed
// (k ascending, the three products in the order above), no atomics, no split-K: the same inputs
// give the same bits in the encoder and in the decoder.
//
// Tile: 128 rows (UMMA_M = 128, cta_group::1) x up to 256 columns (UMMA_N = tile width) x 64 K per stage.
// Warp roles (192 threads):
//   warps 0-3  A producers: implicit-GEMM gather of 16-byte (8-channel) chunks with cp.async straight
//              into the 128-byte-swizzled K-major shared-memory layout, zero-filling SAME padding,
//              transposed-convolution borders, K tails and M tails; completion is signalled with
//              cp.async.mbarrier.arrive.noinc.  After the main loop the same warps run the epilogue:
//              tcgen05.ld -> bias -> LeakyReLU -> hi/lo split -> 16-byte stores (or the final epilogue).
//   warp 4     B producer: one cp.async.bulk per stage; the weights were pre-tiled on the host as the
//              exact shared-memory image (hi plane then lo plane), so no tensor map is needed.
//   warp 5     TMEM allocation + single-thread MMA issue; tcgen05.commit frees the stage / publishes
//              the accumulator.
#include "kernels_common.cuh"

namespace pnn {

namespace {

constexpr int STAGES = 2;
constexpr int A_PLANE = TC_BM * 128;                 // 16 KB
constexpr int B_PLANE_MAX = TC_BN * 128;             // 32 KB
constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE_MAX;   // 96 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start address >> 4 in
// bits [0,14), leading byte offset (unused for swizzled K-major, 1) in [16,30), stride byte offset
// = 1024 B between 8-row groups in [32,46), descriptor version 1 in [46,48), layout SWIZZLE_128B = 2
// in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(GemmLaunch L) {
    extern __shared__ uint8_t smem_raw[];
    const GemmGeom& g = L.g;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
    // barrier slots (8 B each): full_a[s], full_b[s], empty[s], tmem_full; then the TMEM base address
    auto full_a = [&](int s) { return bars + 8u * s; };
    auto full_b = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    const uint32_t tmem_full = bars + 8u * (3 * STAGES);
    const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_nt = (g.N + TC_BN - 1) / TC_BN;
    const int nt = blockIdx.x % num_nt, mt = blockIdx.x / num_nt;
    const int m0 = mt * TC_BM, n0 = nt * TC_BN;
    int bn = g.N - n0;
    bn = bn > TC_BN ? TC_BN : ((bn + 15) & ~15);
    const int num_kb = (g.K + TC_BK - 1) / TC_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_a(s), 128);
            mbar_init(full_b(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp < 4) {
        // ------------------------------------------------------------------ A producers
        const int c = threadIdx.x & 7;          // 16-byte chunk (8 channels) inside the 64-wide k block
        const int rgrp = threadIdx.x >> 3;      // rows rgrp + 16*i
        const uint32_t sw = (uint32_t)((c ^ (rgrp & 7)) << 4);
        int64_t rbase[8];
        int rpos[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + rgrp + 16 * i;
            if (m < L.M) {
                const int b = m / g.P, p = m - b * g.P;
                const int oy = p / g.OW, ox = p - oy * g.OW;
                rbase[i] = (int64_t)b * g.in_sample_stride;
                rpos[i] = (oy << 16) | ox;
            } else {
                rbase[i] = 0;
                rpos[i] = -1;
            }
        }
        const __nv_bfloat16* in_hi = (const __nv_bfloat16*)L.in.p0;
        const __nv_bfloat16* in_lo = (const __nv_bfloat16*)L.in.p1;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t round = (uint32_t)(kb / STAGES);
            mbar_wait(empty(s), (round & 1u) ^ 1u);
            const int k = kb * TC_BK + c * 8;
            const bool k_ok = k < g.K;
            const int tap = k / g.Cin, ci = k - tap * g.Cin;
            const int ty = tap / g.TW, tx = tap - ty * g.TW;
            const int dy = ty * g.sy_t + g.cy, dx = tx * g.sx_t + g.cx;
            const uint32_t a_hi = smem_base + s * STAGE_BYTES + sw;
            const uint32_t a_lo = a_hi + A_PLANE;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int oy = rpos[i] >> 16, ox = rpos[i] & 0xffff;
                const int iy = oy * g.sy_o + dy, ix = ox * g.sx_o + dx;
                const bool ok = k_ok && rpos[i] >= 0 && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
                const int64_t off = ok ? rbase[i] + ((int64_t)iy * g.IW + ix) * g.Cin + ci : 0;
                const uint32_t row_off = (uint32_t)(rgrp + 16 * i) * 128u;
                cp_async_16(a_hi + row_off, in_hi + off, ok ? 16u : 0u);
                cp_async_16(a_lo + row_off, in_lo + off, ok ? 16u : 0u);
            }
            cp_async_arrive_noinc(full_a(s));
        }

        // ------------------------------------------------------------------ epilogue
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = warp * 32 + lane;
        const int m = m0 + row;
        const bool row_ok = m < L.M;
        int64_t obase = 0;
        if (row_ok) {
            const int b = m / g.P, p = m - b * g.P;
            const int oy = p / g.OW, ox = p - oy * g.OW;
            obase = (int64_t)b * g.out_sample_stride +
                    ((int64_t)(oy * g.osy + g.ooy) * g.OWf + (ox * g.osx + g.oox)) * g.N;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int cg = 0; cg < bn / 16; ++cg) {
            uint32_t v[16];
            tmem_ld16(taddr + cg * 16, v);
            const int n = n0 + cg * 16;
            if (!row_ok || n >= g.N) continue;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                f[j] = __uint_as_float(v[j]) + __ldg(L.bias + n + j);
                if (g.leaky) f[j] = leaky_relu(f[j]);
            }
            if (L.out_mode == OUT_FINAL) {
#pragma unroll
                for (int j = 0; j < 16; ++j) final_store(L.fin, obase + n + j, f[j]);
            } else {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(f[2 * j], h0, l0);
                    split_bf16(f[2 * j + 1], h1, l1);
                    hi[j] = pack_bf16(h0, h1);
                    lo[j] = pack_bf16(l0, l1);
                }
                uint4* ph = (uint4*)((__nv_bfloat16*)L.out.p0 + obase + n);
                uint4* pl = (uint4*)((__nv_bfloat16*)L.out.p1 + obase + n);
                ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (warp == 4) {
        // ------------------------------------------------------------------ B producer
        if (lane == 0) {
            const uint32_t stage_b_bytes = 2u * (uint32_t)bn * 128u;
            const uint8_t* src = L.w_tiles + (size_t)nt * num_kb * (2u * TC_BN * 128u);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t round = (uint32_t)(kb / STAGES);
                mbar_wait(empty(s), (round & 1u) ^ 1u);
                mbar_arrive_expect_tx(full_b(s), stage_b_bytes);
                bulk_copy_g2s(smem_base + s * STAGE_BYTES + 2 * A_PLANE, src + (size_t)kb * stage_b_bytes, stage_b_bytes,
                              full_b(s));
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer
        // instruction descriptor: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major,
        // N >> 3 in [17,23), M >> 4 in [24,29)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t round = (uint32_t)(kb / STAGES);
            mbar_wait(full_a(s), round & 1u);
            mbar_wait(full_b(s), round & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                // cp.async (generic proxy) wrote the A tiles; order them before the tensor core's
                // async-proxy reads
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const uint32_t a_hi = smem_base + s * STAGE_BYTES;
                const uint32_t a_lo = a_hi + A_PLANE;
                const uint32_t b_hi = a_hi + 2 * A_PLANE;
                const uint32_t b_lo = b_hi + (uint32_t)bn * 128u;
#pragma unroll
                for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
                    const uint64_t da_hi = make_desc(a_hi + k4 * 32), da_lo = make_desc(a_lo + k4 * 32);
                    const uint64_t db_hi = make_desc(b_hi + k4 * 32), db_lo = make_desc(b_lo + k4 * 32);
                    umma_bf16(tmem_base, da_hi, db_hi, idesc, (kb | k4) != 0 ? 1u : 0u);
                    umma_bf16(tmem_base, da_hi, db_lo, idesc, 1u);
                    umma_bf16(tmem_base, da_lo, db_hi, idesc, 1u);
                }
                umma_commit(empty(s));
                if (kb == num_kb - 1) umma_commit(tmem_full);
            }
            __syncwarp();
        }
    }

    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace

cudaError_t gemm_tc_init() {
    return cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
}

int launch_gemm_tc(const GemmLaunch& L, cudaStream_t stream) {
    if (L.M == 0) return 0;
    const int num_nt = (L.g.N + TC_BN - 1) / TC_BN;
    const int num_mt = (L.M + TC_BM - 1) / TC_BM;
    gemm_tc_kernel<<<num_nt * num_mt, NUM_THREADS, SMEM_BYTES, stream>>>(L);
    return 1;
}

}  // namespace pnn
