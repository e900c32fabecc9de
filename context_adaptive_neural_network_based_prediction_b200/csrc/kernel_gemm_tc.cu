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
// Persistent kernel: one CTA per SM, tiles (128 rows x up to 256 columns) taken round-robin
// (tile = blockIdx.x + i*gridDim.x, n-tile fastest so that concurrent CTAs share the A rows in L2).
// Warp roles (320 threads):
//   warps 0-3  A producers.  Real convolutions (more than one tap, Cin % 64 == 0): the M tile is a box of
//              bw x bh output pixels x nb samples and ONE thread issues two cp.async.bulk.tensor (TMA, rank-4
//              NHWC tensor map, hardware 128B swizzle, out-of-bounds zero fill = SAME padding / tconv borders,
//              element strides for stride 2) per K block.  Other layers (FC, 1x1, K tails): implicit-GEMM
//              gather of 16-byte (8-channel) chunks with cp.async straight into the 128-byte-swizzled K-major
//              layout, zero-filling K and M tails; completion via cp.async.mbarrier.arrive.noinc.
//   warp 4     B producer: one cp.async.bulk per stage; the weights were pre-tiled on the host as the
//              exact shared-memory image (hi plane then lo plane), so no tensor map is needed.
//   warp 5     TMEM allocation (512 columns = two accumulators) + single-thread MMA issue;
//              tcgen05.commit frees the smem stage / publishes the accumulator.
//   warps 8-11 / 12-15  two epilogue sets, one per accumulator, so that the epilogues of tiles i and i+1
//              and the main loop of tile i+2 overlap: tcgen05.ld -> bias (warp-uniform 16-byte loads) -> LeakyReLU ->
//              hi/lo split (packed cvt) -> 32-byte global stores straight from registers, a lane owning the 64 contiguous
//              bytes per plane of its row in a 32-column block (or the final epilogue: raw fp32 + add-mean / clip / round).
//   (warps 6-7 idle: they keep the epilogue sets aligned on warpgroups / TMEM lane quarters)
// The 224 KB smem ring has 2-6 stages depending on the tile width (A 32 KB + B 2*bn*128 B per stage); in tap-reuse mode
// it is split into an A ring (2-3 x 40 KB) and a B ring (4-8 stages).
#include "kernels_common.cuh"

#include <cuda.h>

#include <cstdlib>

namespace pnn {

namespace {

constexpr int MAX_STAGES = 8;
constexpr int A_PLANE = TC_BM * 128;                 // 16 KB
constexpr int RING_BYTES = 224 * 1024;               // 2 stages at bn = 256 (96 KB each), 3 at 128, 4 at 64
constexpr int SMEM_BYTES = RING_BYTES + 1024 /*alignment slack*/ + 512 /*barriers*/;
constexpr int XR_A_ROWS = 160;                       // tap-reuse box: (8 + 2) x bh x nb rows
constexpr int XR_A_PLANE = XR_A_ROWS * 128;          // 20 KB
constexpr int NUM_THREADS = 512;
constexpr int TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp.  The single-thread roles (TMA / bulk-copy producers, MMA issuer) run their whole loop
// under this predicate: the compiler then knows exactly one thread is active and emits the uniform-datapath
// instructions (UTCHMMA, UTMALDG, UBLKCP, UTCBAR) directly instead of an election loop with vector-to-uniform
// register moves around every one of them, which made the issuer instruction-bound (~0.66 us per K block).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}

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
// K-major SWIZZLE_128B descriptor = {lo: start address >> 4 | LBO 1 << 16, hi: SBO (1024 >> 4) | version 1 << 14 | SWIZZLE_128B 2 << 29}
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
        : "memory");
}
// same with its own high word for the A descriptor (stride between 8-row groups other than 1024 bytes)
__device__ __forceinline__ void umma_bf16_x(uint32_t tmem_d, uint32_t adesc_lo, uint32_t adesc_hi, uint32_t bdesc_lo, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %6};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI), "r"(adesc_hi)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// Row index m -> (sample b, oy, ox) of this launch's output positions, advanced by a constant step without
// divisions: one unsigned division pair when a tile starts, then adds and compares (step <= P, so at most
// one wrap per level).  FC layers (P == 1) have b = m.
struct RowIter {
    unsigned b, oy, ox;
};
struct RowStep {
    unsigned P, OW, OH, sdiv, smod, step;
};
__device__ __forceinline__ RowStep make_row_step(const GemmGeom& g, unsigned step) {
    RowStep r;
    r.P = (unsigned)g.P;
    r.OW = (unsigned)g.OW;
    r.OH = r.P / r.OW;
    r.sdiv = step / r.OW;
    r.smod = step - r.sdiv * r.OW;
    r.step = step;
    return r;
}
__device__ __forceinline__ RowIter row_init(const RowStep& rs, unsigned m) {
    RowIter it;
    if (rs.P == 1u) {
        it.b = m; it.oy = 0; it.ox = 0;
    } else {
        it.b = m / rs.P;
        const unsigned p = m - it.b * rs.P;
        it.oy = p / rs.OW;
        it.ox = p - it.oy * rs.OW;
    }
    return it;
}
__device__ __forceinline__ void row_advance(const RowStep& rs, RowIter& it) {
    if (rs.P == 1u) {
        it.b += rs.step;
    } else {
        it.ox += rs.smod;
        it.oy += rs.sdiv;
        if (it.ox >= rs.OW) { it.ox -= rs.OW; it.oy += 1; }
        if (it.oy >= rs.OH) { it.oy -= rs.OH; it.b += 1; }
    }
}
__device__ __forceinline__ int64_t out_offset(const GemmGeom& g, const RowIter& it) {
    return (int64_t)it.b * g.out_sample_stride +
           ((int64_t)((int)it.oy * g.osy + g.ooy) * g.OWf + ((int)it.ox * g.osx + g.oox)) * g.N;
}

__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(GemmLaunch L, const __grid_constant__ CUtensorMap tmap_hi,
                                                                   const __grid_constant__ CUtensorMap tmap_lo) {
    extern __shared__ uint8_t smem_raw[];
    const GemmGeom& g = L.g;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + RING_BYTES;
    // barrier slots (8 B each)
    auto full_a = [&](int s) { return bars + 8u * s; };
    auto full_b = [&](int s) { return bars + 8u * (MAX_STAGES + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * MAX_STAGES + s); };
    auto tmem_full = [&](int a) { return bars + 8u * (3 * MAX_STAGES + a); };
    auto tmem_empty = [&](int a) { return bars + 8u * (3 * MAX_STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (3 * MAX_STAGES + 4);
    auto empty_a = [&](int s) { return bars + 8u * (3 * MAX_STAGES + 5 + s); };   // tap-reuse mode: A ring of its own

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_nt = (g.N + TC_BN - 1) / TC_BN;
    // TMA mode: an M tile is a box of (1 << bw_log2) x (1 << bh_log2) output pixels x nb samples
    const int box_shift = L.bw_log2 + L.bh_log2;
    const int nb_box = TC_BM >> box_shift;
    const int n_samples = L.M / g.P;
    const int num_mt = L.tma ? ((n_samples + nb_box - 1) / nb_box) * L.y_tiles * L.x_tiles : (L.M + TC_BM - 1) / TC_BM;
    const int num_kb = (g.K + TC_BK - 1) / TC_BK;
    // split-K: tile = (ks, mt, nt) with nt fastest; slice ks owns the K blocks [ks*kb_per, min(num_kb, (ks+1)*kb_per))
    const int split_k = L.split_k > 1 ? L.split_k : 1;
    const int kb_per = (num_kb + split_k - 1) / split_k;
    const int unit = blockIdx.x, units = gridDim.x;
    const int mn_tiles = num_nt * num_mt;
    const int num_tiles = mn_tiles * split_k;
    // the widest tile fixes the stage size, hence the ring depth
    const int bn_max = g.N >= TC_BN ? TC_BN : ((g.N + 15) & ~15);
    const uint32_t stage_bytes = 2u * A_PLANE + 2u * (uint32_t)bn_max * 128u;
    int num_stages = RING_BYTES / stage_bytes;
    if (num_stages > MAX_STAGES) num_stages = MAX_STAGES;
    // Narrow layers (N <= 128): the hi and lo weight tiles are contiguous in shared memory, so A_hi * [B_hi ; B_lo] is
    // ONE MMA of width 2*bn whose two column halves (hi*hi and hi*lo) are added in the epilogue; with A_lo * B_hi that
    // makes 2 MMAs per K step instead of 3 and a third less A traffic from shared memory.
    const bool merged = g.N <= 128;
    // tap-reuse mode: A ring (stages of two 20 KB planes) then B ring
    // (the B ring must be deep: a K block of a narrow layer is ~0.35 us of tensor work against a ~2 us load round trip)
    const int xr_na = bn_max <= 64 ? 3 : 2, xr_nb = bn_max <= 32 ? 8 : (bn_max <= 64 ? 6 : 4);
    const uint32_t xr_b_bytes = 2u * (uint32_t)bn_max * 128u;
    const uint32_t xr_b_base = smem_base + (uint32_t)xr_na * 2u * XR_A_PLANE;
    const int chunks = g.Cin >> 6;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            mbar_init(full_a(s), L.tma ? 1 : 128);
            mbar_init(full_b(s), 1);
            mbar_init(empty(s), 1);
            mbar_init(empty_a(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full(a), 1);
            mbar_init(tmem_empty(a), 4);
        }
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

    if (warp < 4 && L.xr) {
        // ------------------------------------------------------------------ A producer, tap reuse: one box per (ty, chunk, group)
        if (warp == 0 && elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_lo) : "memory");
            int s = 0;
            uint32_t ph = 0;
            for (int tile = unit; tile < num_tiles; tile += units) {
                int r = tile / num_nt;
                const int xt = r % L.x_tiles;
                r /= L.x_tiles;
                const int yt = r % L.y_tiles, bt = r / L.y_tiles;
                const int ox0 = xt << L.bw_log2, oy0 = yt << L.bh_log2;
                for (int ty = 0; ty < g.TH; ++ty) {
                    const int c2 = oy0 * g.sy_o + ty * g.sy_t + g.cy;
                    for (int ch = 0; ch < chunks; ++ch) {
                        for (int xg = 0; xg < L.xr_groups; ++xg) {
                            mbar_wait(empty_a(s), ph ^ 1u);
                            const uint32_t a_hi = smem_base + (uint32_t)s * 2u * XR_A_PLANE;
                            mbar_arrive_expect_tx(full_a(s), 2u * XR_A_PLANE);
                            tma_load_4d(a_hi, &tmap_hi, ch * 64, ox0 * g.sx_o + L.xr_start[xg], c2, bt * nb_box, full_a(s));
                            tma_load_4d(a_hi + XR_A_PLANE, &tmap_lo, ch * 64, ox0 * g.sx_o + L.xr_start[xg], c2, bt * nb_box, full_a(s));
                            if (++s == xr_na) { s = 0; ph ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp < 4 && L.tma) {
        // ------------------------------------------------------------------ A producer, TMA: one thread, two box loads per K block
        if (warp == 0 && elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_lo) : "memory");
            int s = 0;
            uint32_t ph = 0;
            for (int tile = unit; tile < num_tiles; tile += units) {
                const int ks = tile / mn_tiles;
                int r = (tile - ks * mn_tiles) / num_nt;
                const int xt = r % L.x_tiles;
                r /= L.x_tiles;
                const int yt = r % L.y_tiles, bt = r / L.y_tiles;
                const int kb0 = ks * kb_per, kb1 = min(num_kb, kb0 + kb_per);
                const int ox0 = xt << L.bw_log2, oy0 = yt << L.bh_log2;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty(s), ph ^ 1u);
                    const int k = kb * TC_BK;
                    const int tap = k / g.Cin, ci = k - tap * g.Cin;
                    const int ty = tap / g.TW, tx = tap - ty * g.TW;
                    const int c1 = ox0 * g.sx_o + tx * g.sx_t + g.cx, c2 = oy0 * g.sy_o + ty * g.sy_t + g.cy;
                    const uint32_t a_hi = smem_base + s * stage_bytes;
                    if (L.debug_flags & 1) {
                        mbar_arrive(full_a(s));
                    } else {
                        mbar_arrive_expect_tx(full_a(s), 2u * A_PLANE);
                        tma_load_4d(a_hi, &tmap_hi, ci, c1, c2, bt * nb_box, full_a(s));
                        tma_load_4d(a_hi + A_PLANE, &tmap_lo, ci, c1, c2, bt * nb_box, full_a(s));
                    }
                    if (++s == num_stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp < 4) {
        // ------------------------------------------------------------------ A producers
        const int c = threadIdx.x & 7;          // 16-byte chunk (8 channels) inside the 64-wide k block
        const int rgrp = threadIdx.x >> 3;      // rows rgrp + 16*i
        const uint32_t sw = (uint32_t)((c ^ (rgrp & 7)) << 4);
        const __nv_bfloat16* in_hi = (const __nv_bfloat16*)L.in.p0;
        const __nv_bfloat16* in_lo = (const __nv_bfloat16*)L.in.p1;
        const RowStep rs16 = make_row_step(g, 16u);
        int s = 0;
        uint32_t ph = 0;
        int cur_mt = -1;
        int64_t rbase[8];
        int rpos[8];
        for (int tile = unit; tile < num_tiles; tile += units) {
            const int ks = tile / mn_tiles;
            const int mt = (tile - ks * mn_tiles) / num_nt;
            const int kb0 = ks * kb_per, kb1 = min(num_kb, kb0 + kb_per);
            if (mt != cur_mt) {
                cur_mt = mt;
                RowIter it = row_init(rs16, (unsigned)(mt * TC_BM + rgrp));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int m = mt * TC_BM + rgrp + 16 * i;
                    if (m < L.M) {
                        rbase[i] = (int64_t)it.b * g.in_sample_stride;
                        rpos[i] = (int)((it.oy << 16) | it.ox);
                    } else {
                        rbase[i] = 0;
                        rpos[i] = -1;
                    }
                    row_advance(rs16, it);
                }
            }
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty(s), ph ^ 1u);
                const int k = kb * TC_BK + c * 8;
                const bool k_ok = k < g.K;
                const int tap = k / g.Cin, ci = k - tap * g.Cin;
                const int ty = tap / g.TW, tx = tap - ty * g.TW;
                const int dy = ty * g.sy_t + g.cy, dx = tx * g.sx_t + g.cx;
                const uint32_t a_hi = smem_base + s * stage_bytes + sw;
                const uint32_t a_lo = a_hi + A_PLANE;
                if (!(L.debug_flags & 1)) {
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
                }
                cp_async_arrive_noinc(full_a(s));
                if (++s == num_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 4 && L.xr) {
        // ------------------------------------------------------------------ B producer, tap reuse: K blocks in (ty, chunk, group, tap) order
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = unit; tile < num_tiles; tile += units) {
                const int nt = tile % num_nt;
                int bn = g.N - nt * TC_BN;
                bn = bn > TC_BN ? TC_BN : ((bn + 15) & ~15);
                const uint32_t b_bytes = 2u * (uint32_t)bn * 128u;
                const uint8_t* src = L.w_tiles + (size_t)nt * num_kb * (2u * TC_BN * 128u);
                for (int ty = 0; ty < g.TH; ++ty)
                    for (int ch = 0; ch < chunks; ++ch)
                        for (int xg = 0; xg < L.xr_groups; ++xg)
                            for (int t = 0; t < L.xr_ntaps[xg]; ++t) {
                                const int kb = (ty * g.TW + L.xr_tx[xg][t]) * chunks + ch;
                                mbar_wait(empty(s), ph ^ 1u);
                                mbar_arrive_expect_tx(full_b(s), b_bytes);
                                bulk_copy_g2s(xr_b_base + (uint32_t)s * xr_b_bytes, src + (size_t)kb * b_bytes, b_bytes, full_b(s));
                                if (++s == xr_nb) { s = 0; ph ^= 1u; }
                            }
            }
        }
    } else if (warp == 4) {
        // ------------------------------------------------------------------ B producer
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = unit; tile < num_tiles; tile += units) {
                const int ks = tile / mn_tiles;
                const int nt = (tile - ks * mn_tiles) % num_nt;
                const int kb0 = ks * kb_per, kb1 = min(num_kb, kb0 + kb_per);
                int bn = g.N - nt * TC_BN;
                bn = bn > TC_BN ? TC_BN : ((bn + 15) & ~15);
                const uint32_t b_bytes = 2u * (uint32_t)bn * 128u;
                const uint8_t* src = L.w_tiles + (size_t)nt * num_kb * (2u * TC_BN * 128u);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty(s), ph ^ 1u);
                    if (L.debug_flags & 2) {
                        mbar_arrive(full_b(s));
                    } else {
                        mbar_arrive_expect_tx(full_b(s), b_bytes);
                        bulk_copy_g2s(smem_base + s * stage_bytes + 2 * A_PLANE, src + (size_t)kb * b_bytes, b_bytes, full_b(s));
                    }
                    if (++s == num_stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 5 && L.xr) {
        // ------------------------------------------------------------------ MMA issuer, tap reuse (one elected thread runs the loop)
        if (elect_one()) {
            int sa = 0, sb = 0;
            uint32_t pha = 0, phb = 0;
            int lt = 0;
            // A descriptor high word: 10 rows of 128 bytes between the 8-row groups
            constexpr uint32_t A_DESC_HI = (1280u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t da_ring = desc_lo(smem_base), db_ring = desc_lo(xr_b_base);
            for (int tile = unit; tile < num_tiles; tile += units, ++lt) {
                const int nt = tile % num_nt;
                int bn = g.N - nt * TC_BN;
                bn = bn > TC_BN ? TC_BN : ((bn + 15) & ~15);
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
                const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * bn) >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
                const uint32_t b_lo_step = ((uint32_t)bn * 128u) >> 4;
                const int ab = lt & 1;
                const uint32_t acc = tmem_base + (uint32_t)(ab * TC_BN);
                mbar_wait(tmem_empty(ab), (((uint32_t)lt >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t accumulate = 0;
                for (int ty = 0; ty < g.TH; ++ty) {
                    for (int ch = 0; ch < chunks; ++ch) {
                        for (int xg = 0; xg < L.xr_groups; ++xg) {
                            mbar_wait(full_a(sa), pha);
                            const int ntaps = L.xr_ntaps[xg];
                            const bool last_group = ty == g.TH - 1 && ch == chunks - 1 && xg == L.xr_groups - 1;
                            const uint32_t da_stage = da_ring + (uint32_t)sa * ((2u * XR_A_PLANE) >> 4);
                            for (int t = 0; t < ntaps; ++t) {
                                mbar_wait(full_b(sb), phb);
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                                const uint32_t da_hi = da_stage + (uint32_t)L.xr_off[xg][t] * (128u >> 4);
                                const uint32_t da_lo = da_hi + (XR_A_PLANE >> 4);
                                const uint32_t db_hi = db_ring + (uint32_t)sb * (xr_b_bytes >> 4);
                                const uint32_t db_lo = db_hi + b_lo_step;
                                if (merged) {
#pragma unroll
                                    for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
                                        umma_bf16_x(acc, da_hi + 2 * k4, A_DESC_HI, db_hi + 2 * k4, idesc2, accumulate);   // [hi*hi | hi*lo]
                                        umma_bf16_x(acc, da_lo + 2 * k4, A_DESC_HI, db_hi + 2 * k4, idesc, 1u);            // lo*hi
                                        accumulate = 1u;
                                    }
                                } else {
#pragma unroll
                                    for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
                                        umma_bf16_x(acc, da_hi + 2 * k4, A_DESC_HI, db_hi + 2 * k4, idesc, accumulate);
                                        umma_bf16_x(acc, da_hi + 2 * k4, A_DESC_HI, db_lo + 2 * k4, idesc, 1u);
                                        umma_bf16_x(acc, da_lo + 2 * k4, A_DESC_HI, db_hi + 2 * k4, idesc, 1u);
                                        accumulate = 1u;
                                    }
                                }
                                umma_commit(empty(sb));
                                if (t == ntaps - 1) {
                                    umma_commit(empty_a(sa));
                                    if (last_group) umma_commit(tmem_full(ab));
                                }
                                if (++sb == xr_nb) { sb = 0; phb ^= 1u; }
                            }
                            if (++sa == xr_na) { sa = 0; pha ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issuer (one elected thread runs the loop)
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            int lt = 0;
            const uint32_t d_ring = desc_lo(smem_base), d_stage = stage_bytes >> 4;
            // 16-wide K steps of the last K block that hold data (the K tail is skipped, not multiplied by zero)
            const int last_ksteps = (g.K - (num_kb - 1) * TC_BK + 15) >> 4;
            const bool needs_proxy_fence = !L.tma && !(L.debug_flags & 8);   // cp.async (generic proxy) wrote the A tiles
            for (int tile = unit; tile < num_tiles; tile += units, ++lt) {
                const int ks = tile / mn_tiles;
                const int nt = (tile - ks * mn_tiles) % num_nt;
                const int kb0 = ks * kb_per, kb1 = min(num_kb, kb0 + kb_per);
                int bn = g.N - nt * TC_BN;
                bn = bn > TC_BN ? TC_BN : ((bn + 15) & ~15);
                // instruction descriptor: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major,
                // N >> 3 in [17,23), M >> 4 in [24,29)
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
                const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * bn) >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
                const uint32_t b_lo_step = ((uint32_t)bn * 128u) >> 4;
                const int ab = lt & 1;
                const uint32_t acc = tmem_base + (uint32_t)(ab * TC_BN);
                mbar_wait(tmem_empty(ab), (((uint32_t)lt >> 1) & 1u) ^ 1u);     // the epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t accumulate = 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (!(L.debug_flags & 64)) {
                        mbar_wait(full_a(s), ph);
                        mbar_wait(full_b(s), ph);
                    }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // order the generic-proxy writes of the A tiles before the tensor core's async-proxy reads
                    if (needs_proxy_fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    const uint32_t da_hi = d_ring + (uint32_t)s * d_stage;
                    const uint32_t da_lo = da_hi + (A_PLANE >> 4);
                    const uint32_t db_hi = da_hi + ((2 * A_PLANE) >> 4);
                    const uint32_t db_lo = db_hi + b_lo_step;
                    int ksteps = kb == num_kb - 1 ? last_ksteps : TC_BK / 16;
                    if (L.debug_flags & 16) ksteps = 1;
                    if (merged) {
#pragma unroll
                        for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
                            if (k4 < ksteps) {                  // +2 per K step: 32 bytes >> 4 inside the 128-byte swizzle row
                                umma_bf16(acc, da_hi + 2 * k4, db_hi + 2 * k4, idesc2, accumulate);   // [hi*hi | hi*lo]
                                umma_bf16(acc, da_lo + 2 * k4, db_hi + 2 * k4, idesc, 1u);            // lo*hi
                                accumulate = 1u;
                            }
                        }
                    } else {
#pragma unroll
                        for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
                            if (k4 < ksteps) {
                                umma_bf16(acc, da_hi + 2 * k4, db_hi + 2 * k4, idesc, accumulate);
                                umma_bf16(acc, da_hi + 2 * k4, db_lo + 2 * k4, idesc, 1u);
                                umma_bf16(acc, da_lo + 2 * k4, db_hi + 2 * k4, idesc, 1u);
                                accumulate = 1u;
                            }
                        }
                    }
                    if (L.debug_flags & 32) mbar_arrive(empty(s));
                    else umma_commit(empty(s));
                    if (kb == kb1 - 1) umma_commit(tmem_full(ab));
                    if (++s == num_stages) { s = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ epilogue warps (set 0: 8..11, set 1: 12..15)
        const int set = (warp - 8) >> 2;                     // = accumulator buffer this set drains
        const int q = warp & 3;                              // TMEM lane quarter this warp may read
        const RowStep rs1 = make_row_step(g, 1u);
        int j = 0;                                           // tiles this set has processed
        for (int tile = unit + set * units; tile < num_tiles; tile += 2 * units, ++j) {
            const int ks = tile / mn_tiles;
            const int nt = (tile - ks * mn_tiles) % num_nt, mt = (tile - ks * mn_tiles) / num_nt;
            const int n0 = nt * TC_BN;
            int bn = g.N - n0;
            bn = bn > TC_BN ? TC_BN : ((bn + 15) & ~15);
            int m_own = mt * TC_BM + q * 32 + lane;           // accumulator row owned in the TMEM-read phase
            // every lane stores the row it reads from TMEM
            int64_t obase_own = -1;
            if (L.tma) {
                // box order: row r of the tile = (sample r >> box_shift, y (r >> bw_log2) & (bh - 1), x r & (bw - 1))
                int rr = mt;
                const int xt = rr % L.x_tiles;
                rr /= L.x_tiles;
                const int yt = rr % L.y_tiles, bt = rr / L.y_tiles;
                const int bw_mask = (1 << L.bw_log2) - 1, bh_mask = (1 << L.bh_log2) - 1;
                const int r = q * 32 + lane;
                RowIter it;
                it.b = (unsigned)(bt * nb_box + (r >> box_shift));
                it.oy = (unsigned)((yt << L.bh_log2) + ((r >> L.bw_log2) & bh_mask));
                it.ox = (unsigned)((xt << L.bw_log2) + (r & bw_mask));
                const bool ok = (int)it.b < n_samples && it.ox < (unsigned)g.OW && it.oy * (unsigned)g.OW < (unsigned)g.P;
                m_own = ok ? (int)(it.b * (unsigned)g.P + it.oy * (unsigned)g.OW + it.ox) : L.M;
                if (ok) obase_own = out_offset(g, it);
            } else if (m_own < L.M) {
                obase_own = out_offset(g, row_init(rs1, (unsigned)m_own));
            }
            mbar_wait(tmem_full(set), (uint32_t)j & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * TC_BN);
            for (int c0 = 0; c0 < bn; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                if (merged) {                                 // second half of the accumulator: the hi*lo products
                    uint32_t v2[32];
                    tmem_ld32(taddr + bn + c0, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) v[jj] = __float_as_uint(__uint_as_float(v[jj]) + __uint_as_float(v2[jj]));
                }
                const int nbase = n0 + c0;
                int nvalid = g.N - nbase;                    // multiple of 16 by construction
                if (nvalid > 32) nvalid = 32;
                // the 32 (or 16) bias values of this block: warp-uniform 16-byte loads (one L1 broadcast each) instead of a
                // shuffle per output element
                float bias[32];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float4 b4 = jj * 4 < nvalid ? __ldg((const float4*)(L.bias + nbase) + jj) : make_float4(0.f, 0.f, 0.f, 0.f);
                    bias[4 * jj] = b4.x; bias[4 * jj + 1] = b4.y; bias[4 * jj + 2] = b4.z; bias[4 * jj + 3] = b4.w;
                }
                tmem_ld_wait();
                if (L.debug_flags & 4) continue;
                if (split_k > 1) {
                    // raw accumulators of this K slice; bias / activation / split happen in splitk_reduce_kernel
                    if (m_own < L.M) {
                        float4* dst = (float4*)(L.partial + ((size_t)ks * L.M + m_own) * g.N + nbase);
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            if (jj * 4 < nvalid) {
                                dst[jj] = make_float4(__uint_as_float(v[4 * jj]), __uint_as_float(v[4 * jj + 1]),
                                                      __uint_as_float(v[4 * jj + 2]), __uint_as_float(v[4 * jj + 3]));
                            }
                        }
                    }
                    continue;
                }
                if (L.out_mode == OUT_FINAL) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        if (jj < nvalid && obase_own >= 0) {
                            float f = __uint_as_float(v[jj]) + bias[jj];
                            if (g.leaky) f = leaky_relu(f);
                            final_store(L.fin, obase_own + nbase + jj, f);
                        }
                    }
                    continue;
                }
                // bias, LeakyReLU, hi/lo split, and straight to global memory: a lane owns 64 contiguous bytes per plane of its
                // row in this block, written as 32-byte sectors (st.global.v8; rows start on multiples of 32 bytes because N
                // is a multiple of 16).  An earlier version staged the tile in shared memory to get 64-byte runs per 4 lanes;
                // with 256-bit stores that costs more issue slots than it saves, and these warps are issue bound.
                if (obase_own >= 0) {
                    __nv_bfloat16* out_hi = (__nv_bfloat16*)L.out.p0 + obase_own + nbase;
                    __nv_bfloat16* out_lo = (__nv_bfloat16*)L.out.p1 + obase_own + nbase;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (half * 16 < nvalid) {
                            uint32_t hi[8], lo[8];
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) {
                                const int col = half * 16 + 2 * jj;
                                float f0 = __uint_as_float(v[col]) + bias[col];
                                float f1 = __uint_as_float(v[col + 1]) + bias[col + 1];
                                if (g.leaky) {
                                    f0 = leaky_relu(f0);
                                    f1 = leaky_relu(f1);
                                }
                                split_bf16x2(f0, f1, hi[jj], lo[jj]);
                            }
                            st_global_v8(out_hi + half * 16, hi);
                            st_global_v8(out_lo + half * 16, lo);
                        }
                    }
                }
            }
            // this warp's TMEM reads of the tile are done: hand the accumulator back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(set));
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// Sums the K slices in ascending order (fixed), then bias, LeakyReLU and the hi/lo split.  Thread = (row, 8 columns).
__global__ void __launch_bounds__(256) splitk_reduce_kernel(GemmLaunch L) {
    const GemmGeom& g = L.g;
    const int groups = g.N >> 3;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= L.M * groups) return;
    const int m = idx / groups, n = (idx - m * groups) << 3;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ks = 0; ks < L.split_k; ++ks) {
        const float4* p = (const float4*)(L.partial + ((size_t)ks * L.M + m) * g.N + n);
        const float4 a = p[0], b = p[1];
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
        acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        float f0 = acc[2 * jj] + L.bias[n + 2 * jj], f1 = acc[2 * jj + 1] + L.bias[n + 2 * jj + 1];
        if (g.leaky) {
            f0 = leaky_relu(f0);
            f1 = leaky_relu(f1);
        }
        split_bf16x2(f0, f1, hi[jj], lo[jj]);
    }
    const RowStep rs = make_row_step(g, 1u);
    const int64_t o = out_offset(g, row_init(rs, (unsigned)m)) + n;
    *(uint4*)((__nv_bfloat16*)L.out.p0 + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *(uint4*)((__nv_bfloat16*)L.out.p1 + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace

int launch_splitk_reduce(const GemmLaunch& L, cudaStream_t stream) {
    const int total = L.M * (L.g.N >> 3);
    splitk_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(L);
    return 1;
}

cudaError_t gemm_tc_init() {
    return cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
}

static int g_num_sms = 0;
static int g_tma_enabled = -1;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;

void gemm_tc_set_tma(int enabled) { g_tma_enabled = enabled ? 1 : 0; }

static int log2_pow2_divisor(int v, int cap) {   // log2 of the largest power of two dividing v, not above cap
    int l = 0;
    while ((v & 1) == 0 && (2 << l) <= cap) {
        v >>= 1;
        ++l;
    }
    return l;
}

// NHWC activation plane [n, IH, IW, Cin] as a rank-4 tensor map with a (64 channels, bw pixels, bh pixels, nb samples) box
static bool make_act_map(CUtensorMap* map, const void* base, const GemmGeom& g, int n, int bw, int bh, int nb, int extra_x = 0) {
    const cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.IW, (cuuint64_t)g.IH, (cuuint64_t)n};
    const cuuint64_t strides[3] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.IW * g.Cin * 2, (cuuint64_t)g.in_sample_stride * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)((bw + extra_x) * g.sx_o), (cuuint32_t)(bh * g.sy_o), (cuuint32_t)nb};
    const cuuint32_t estr[4] = {1, (cuuint32_t)g.sx_o, (cuuint32_t)g.sy_o, 1};
    return g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_gemm_tc(const GemmLaunch& L_in, cudaStream_t stream) {
    if (L_in.M == 0) return 0;
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    if (g_tma_enabled < 0) {
        const char* e = getenv("PNN_TMA");          // default on; PNN_TMA=0 selects the cp.async gather for every layer
        g_tma_enabled = e ? atoi(e) != 0 : 1;
    }
    GemmLaunch L = L_in;
    const GemmGeom& g = L.g;
    alignas(64) CUtensorMap map_hi, map_lo;
    memset(&map_hi, 0, sizeof(map_hi));
    memset(&map_lo, 0, sizeof(map_lo));
    L.tma = 0;
    L.xr = 0;
    // FC geometry (one tap, one position per sample, rows contiguous): a rank-4 map with a 1 x 1 x 128-sample box; the K tail
    // of the last block is the map's out-of-bounds zero fill
    static const int fc_tma_enabled = getenv("PNN_FC_TMA") ? atoi(getenv("PNN_FC_TMA")) != 0 : 1;
    const bool fc_like = fc_tma_enabled && g.P == 1 && g.TH * g.TW == 1 && g.IH == 1 && g.IW == 1 && (g.Cin * 2) % 16 == 0 &&
                         g.in_sample_stride == g.Cin && L.split_k <= 1;
    GemmGeom gm = g;                       // geometry the tensor map is built from
    if (fc_like) gm.sx_o = gm.sy_o = 1;
    // real convolutions with 64-channel K blocks, forward stride (tconv phases walk taps backwards with sx_o = 1: fine)
    if (g_tma_enabled && (fc_like || (g.TH * g.TW > 1 && g.Cin % 64 == 0 && g.sx_o >= 1 && g.sy_o >= 1 && g.sx_o <= 2 && g.sy_o <= 2)) &&
        g.P == (g.P / g.OW) * g.OW && L.M % g.P == 0) {
        if (!g_encode_tiled) {
            cudaDriverEntryPointQueryResult qres;
            void* fn = nullptr;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
                qres == cudaDriverEntryPointSuccess) {
                g_encode_tiled = (EncodeTiledFn)fn;
            }
        }
        const int OH = g.P / g.OW;
        // tap reuse along x: boxes of 8 output columns, every kernel-row group within 2 extra box elements
        static const int xr_enabled = getenv("PNN_XREUSE") ? atoi(getenv("PNN_XREUSE")) != 0 : 1;
        L.xr = 0;
        if (xr_enabled && g.OW % 8 == 0 && g.N <= 128 && L.split_k <= 1 && g.TW <= 5 && !L.debug_flags) {
            bool ok = true;
            int ng = 0;
            for (int rho = 0; rho < g.sx_o && ok; ++rho) {
                int min_t = 1 << 30, cnt = 0;
                for (int tx = 0; tx < g.TW; ++tx) {
                    const int v = tx * g.sx_t;
                    if (((v % g.sx_o) + g.sx_o) % g.sx_o == rho) {
                        min_t = v < min_t ? v : min_t;
                        ++cnt;
                    }
                }
                if (cnt == 0) continue;
                if (cnt > 3) { ok = false; break; }
                int t = 0;
                for (int tx = 0; tx < g.TW; ++tx) {
                    const int v = tx * g.sx_t;
                    if (((v % g.sx_o) + g.sx_o) % g.sx_o != rho) continue;
                    const int off = (v - min_t) / g.sx_o;
                    if (off > 2) ok = false;
                    L.xr_tx[ng][t] = tx;
                    L.xr_off[ng][t] = off;
                    ++t;
                }
                L.xr_start[ng] = min_t + g.cx;
                L.xr_ntaps[ng] = cnt;
                ++ng;
            }
            if (ok && ng >= 1) {
                L.xr = 1;
                L.xr_groups = ng;
            }
        }
        const int bwl = L.xr ? 3 : log2_pow2_divisor(g.OW, TC_BM);
        const int bhl = log2_pow2_divisor(OH, TC_BM >> bwl);
        const int bw = 1 << bwl, bh = 1 << bhl, nb = TC_BM / (bw * bh);
        const int n = L.M / g.P;
        const int extra_x = L.xr ? 2 : 0;
        if (g_encode_tiled && (bw + extra_x) * gm.sx_o <= 256 && bh * gm.sy_o <= 256 &&
            make_act_map(&map_hi, L.in.p0, gm, n, bw, bh, nb, extra_x) && make_act_map(&map_lo, L.in.p1, gm, n, bw, bh, nb, extra_x)) {
            L.tma = 1;
            L.bw_log2 = bwl;
            L.bh_log2 = bhl;
            L.x_tiles = g.OW / bw;
            L.y_tiles = OH / bh;
        } else {
            L.xr = 0;
        }
    }
    const int num_nt = (g.N + TC_BN - 1) / TC_BN;
    long long num_mt = (L.M + TC_BM - 1) / TC_BM;
    if (L.tma) {
        const int nb = TC_BM >> (L.bw_log2 + L.bh_log2);
        num_mt = (long long)((L.M / g.P + nb - 1) / nb) * L.y_tiles * L.x_tiles;
    }
    const long long tiles = (long long)num_nt * num_mt * (L.split_k > 1 ? L.split_k : 1);
    const int grid = (int)(tiles < g_num_sms ? tiles : g_num_sms);   // persistent: one CTA per SM
    gemm_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(L, map_hi, map_lo);
    return 1;
}

}  // namespace pnn
