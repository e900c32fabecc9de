// fp32 FFMA implicit-GEMM kernel: the accuracy yard-stick of the GEMM-shaped layers (FC, convolutions,
// transposed-convolution phases).  Same geometry and reduction structure (k ascending) as the tcgen05
// kernel, plain fp32 arithmetic.  Not the throughput path.
#include "kernels_common.cuh"

#include <cooperative_groups.h>

namespace pnn {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

__global__ void __launch_bounds__(NT) gemm_fp32_kernel(GemmLaunch L) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const GemmGeom& g = L.g;
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // A loader: one row, 4 consecutive k per thread
    const int a_row = tid >> 2, a_kq = (tid & 3) * 4;
    const int am = m0 + a_row;
    const bool a_row_ok = am < L.M;
    int a_oy = 0, a_ox = 0;
    int64_t a_base = 0;
    if (a_row_ok) {
        const int b = am / g.P, p = am - b * g.P;
        a_oy = p / g.OW;
        a_ox = p - a_oy * g.OW;
        a_base = (int64_t)b * g.in_sample_stride;
    }
    // B loader: one k row, 4 consecutive n per thread
    const int b_k = tid >> 4, b_n4 = (tid & 15) * 4;

    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const float* in = (const float*)L.in.p0;
    for (int k0 = 0; k0 < g.K; k0 += BK) {
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k = k0 + a_kq;
        if (a_row_ok && k < g.K) {
            const int tap = k / g.Cin, ci = k - tap * g.Cin;
            const int tyy = tap / g.TW, txx = tap - tyy * g.TW;
            const int iy = a_oy * g.sy_o + tyy * g.sy_t + g.cy;
            const int ix = a_ox * g.sx_o + txx * g.sx_t + g.cx;
            if (iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW) {
                av = *(const float4*)(in + a_base + ((int64_t)iy * g.IW + ix) * g.Cin + ci);
            }
        }
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + b_k < g.K && n0 + b_n4 < g.N) {
            bv = *(const float4*)(L.w_fp32 + (int64_t)(k0 + b_k) * g.N + n0 + b_n4);
        }
        __syncthreads();
        As[a_kq + 0][a_row] = av.x;
        As[a_kq + 1][a_row] = av.y;
        As[a_kq + 2][a_row] = av.z;
        As[a_kq + 3][a_row] = av.w;
        *(float4*)&Bs[b_k][b_n4] = bv;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *(const float4*)&As[kk][ty * 4];
            const float4 b = *(const float4*)&Bs[kk][tx * 4];
            const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }

    const int n = n0 + tx * 4;
    if (n >= g.N) return;
    const float4 bias = *(const float4*)(L.bias + n);
    const float bb[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= L.M) continue;
        const int b = m / g.P, p = m - b * g.P;
        const int oy = p / g.OW, ox = p - oy * g.OW;
        const int64_t o = (int64_t)b * g.out_sample_stride +
                          ((int64_t)(oy * g.osy + g.ooy) * g.OWf + (ox * g.osx + g.oox)) * g.N + n;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = acc[i][j] + bb[j];
            if (g.leaky) v = leaky_relu(v);
            if (L.out_mode == OUT_FINAL) final_store(L.fin, o + j, v);
            else ((float*)L.out.p0)[o + j] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Batch-1 (in-loop) GEMM-shaped layer in fp32: the output map of ONE sample has 16 .. 3072 pixels, far too few rows for the
// 128-row tensor-core tiles, and the layer is bound by streaming its weights once.  CTA = 16 output pixels x 16 output
// channels over the WHOLE K (no split-K across CTAs, no partial sums in global memory, no reduce launch).  Thread
// (ks = t / 16, n = t % 16) accumulates rows k = ks, ks + 16, ... of column n for all 16 pixels: one 4-byte weight load
// (a warp covers two rows of 64 contiguous bytes) and four 16-byte shared-memory loads of the activation tile
// a_s[k][16 pixels] (broadcast to the 16 threads of a k) per 16 FMAs.  The 16 k-slices of an output are then added in
// ascending order.  K is consumed in chunks of SK_CHUNK rows (the implicit-im2col gather of a chunk is staged in shared
// memory, zero for padding / transposed-convolution borders).  Fixed order, fp32 throughout.
// A layer is a chain of dependent L2 round trips (one per chunk), so the chunks of a tile are dealt to the CTAs of a
// THREAD-BLOCK CLUSTER (gridDim.z = cluster size S <= 8, CTA r takes chunks r, r + S, ...): every CTA leaves its 16 x 16
// partial sums in its own shared memory, and CTA 0 of the cluster adds them in ascending rank through distributed shared
// memory -- split-K without a partial-sum buffer in global memory, without a second launch and without atomics.
// ---------------------------------------------------------------------------------------------
constexpr int SK_CHUNK = 512;

__global__ void __launch_bounds__(256) gemm_skinny_kernel(GemmLaunch L) {
    __shared__ __align__(16) float a_s[SK_CHUNK][16];             // 32 KB
    __shared__ float part_s[256];                                  // this CTA's partial sums [pixel][channel]
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)blockIdx.z, n_ranks = (int)gridDim.z;
    const GemmGeom& g = L.g;
    const int t = threadIdx.x;
    const int m0 = blockIdx.x * 16, n0 = blockIdx.y * 16;
    const int ks = t >> 4, n = t & 15;   // (L lives in the constant bank: no copy needed)
    const float* in = (const float*)L.in.p0;
    const float* wcol = L.w_fp32 + n0 + n;
    float acc[16];
#pragma unroll
    for (int p = 0; p < 16; ++p) acc[p] = 0.f;
    // gather role: thread = (pixel gp = t % 16, 16-byte slot q = t / 16 + 16 j): a slot is four consecutive input channels
    // of one tap (Cin is a power of two >= 4 in every layer that comes here)
    const int gp = t & 15, gq0 = t >> 4;
    const int gm = m0 + gp;
    const bool g_ok = gm < L.M;
    int g_oy = 0, g_ox = 0;
    int64_t g_base = 0;
    if (g_ok) {
        const int b = gm / g.P, pp = gm - b * g.P;
        g_oy = pp / g.OW;
        g_ox = pp - g_oy * g.OW;
        g_base = (int64_t)b * g.in_sample_stride;
    }
    const int cin_log2 = 31 - __clz(g.Cin);
    for (int k0 = rank * SK_CHUNK; k0 < g.K; k0 += n_ranks * SK_CHUNK) {
        const int kc = min(SK_CHUNK, g.K - k0);
        // every global load of the chunk is requested up front: the thread's 32 weights (rows ks, ks + 16, ...) and its 8
        // 16-byte pieces of the activation tile; one round trip to L2 per chunk
        float w[SK_CHUNK / 16];
#pragma unroll
        for (int u = 0; u < SK_CHUNK / 16; ++u) {
            // volatile asm, row index clamped: unconditional loads that keep their place before the barrier (left alone, ptxas
            // sinks every load next to its FMAs: 32 dependent L2 round trips per chunk)
            const int kk = min(ks + 16 * u, kc - 1);
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(w[u]) : "l"(wcol + (int64_t)(k0 + kk) * g.N));
        }
        float4 v[SK_CHUNK / 64];
#pragma unroll
        for (int j = 0; j < SK_CHUNK / 64; ++j) {
            const int kk = 4 * (gq0 + 16 * j);
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kk < kc && g_ok) {
                const int k = k0 + kk;
                const int tap = k >> cin_log2, ci = k & (g.Cin - 1);
                const int tyy = tap / g.TW, txx = tap - tyy * g.TW;
                const int iy = g_oy * g.sy_o + tyy * g.sy_t + g.cy;
                const int ix = g_ox * g.sx_o + txx * g.sx_t + g.cx;
                if (iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW) {
                    const float* src = in + g_base + (((int64_t)iy * g.IW + ix) << cin_log2) + ci;
                    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w) : "l"(src));
                }
            }
        }
        __syncthreads();                                          // the previous chunk has been consumed
#pragma unroll
        for (int j = 0; j < SK_CHUNK / 64; ++j) {
            const int kk = 4 * (gq0 + 16 * j);
            if (kk < kc) {
                a_s[kk][gp] = v[j].x;
                a_s[kk + 1][gp] = v[j].y;
                a_s[kk + 2][gp] = v[j].z;
                a_s[kk + 3][gp] = v[j].w;
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < SK_CHUNK / 16; ++u) {
            const int kk = ks + 16 * u;
            if (kk < kc) {
                const float4* row = reinterpret_cast<const float4*>(a_s[kk]);
                const float4 a0 = row[0], a1 = row[1], a2 = row[2], a3 = row[3];
                const float wu = w[u];
                acc[0] = fmaf(a0.x, wu, acc[0]);   acc[1] = fmaf(a0.y, wu, acc[1]);   acc[2] = fmaf(a0.z, wu, acc[2]);   acc[3] = fmaf(a0.w, wu, acc[3]);
                acc[4] = fmaf(a1.x, wu, acc[4]);   acc[5] = fmaf(a1.y, wu, acc[5]);   acc[6] = fmaf(a1.z, wu, acc[6]);   acc[7] = fmaf(a1.w, wu, acc[7]);
                acc[8] = fmaf(a2.x, wu, acc[8]);   acc[9] = fmaf(a2.y, wu, acc[9]);   acc[10] = fmaf(a2.z, wu, acc[10]); acc[11] = fmaf(a2.w, wu, acc[11]);
                acc[12] = fmaf(a3.x, wu, acc[12]); acc[13] = fmaf(a3.y, wu, acc[13]); acc[14] = fmaf(a3.z, wu, acc[14]); acc[15] = fmaf(a3.w, wu, acc[15]);
            }
        }
    }
    // the 16 k-slices of every output, added in ascending order: red[ks][pixel][channel] reuses the activation tile
    __syncthreads();
    float* red = &a_s[0][0];                                      // 16 * 16 * 16 floats = 16 KB
#pragma unroll
    for (int p = 0; p < 16; ++p) red[(ks * 16 + p) * 16 + n] = acc[p];
    __syncthreads();
    const int p = t >> 4;                                         // thread = (pixel p, channel n)
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < 16; ++s) sum += red[(s * 16 + p) * 16 + n];
    if (n_ranks > 1) {
        // partial sums of the cluster's CTAs, added by CTA 0 in ascending rank (distributed shared memory)
        part_s[t] = sum;
        cluster.sync();
        if (rank == 0) {
            sum = 0.f;
            for (int r = 0; r < n_ranks; ++r) sum += cluster.map_shared_rank(part_s, r)[t];
        }
        cluster.sync();                                           // nobody leaves while its shared memory may still be read
        if (rank != 0) return;
    }
    const int m = m0 + p, nn = n0 + n;
    if (m >= L.M || nn >= g.N) return;
    float v = sum + L.bias[nn];
    if (g.leaky) v = leaky_relu(v);
    const int b = m / g.P, pp = m - b * g.P;
    const int oy = pp / g.OW, ox = pp - oy * g.OW;
    const int64_t o = (int64_t)b * g.out_sample_stride + ((int64_t)(oy * g.osy + g.ooy) * g.OWf + (ox * g.osx + g.oox)) * g.N + nn;
    if (L.out_mode == OUT_FINAL) final_store(L.fin, o, v);
    else ((float*)L.out.p0)[o] = v;
}

}  // namespace

int launch_gemm_skinny(const GemmLaunch& L, cudaStream_t stream) {
    if (L.M == 0) return 0;
    const int chunks = (L.g.K + SK_CHUNK - 1) / SK_CHUNK;
    const int ranks = chunks < 8 ? chunks : 8;                    // portable cluster size
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((L.M + 15) / 16, (L.g.N + 15) / 16, ranks);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = ranks;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, gemm_skinny_kernel, L);
    return 1;
}

int launch_gemm_fp32(const GemmLaunch& L, cudaStream_t stream) {
    if (L.M == 0) return 0;
    dim3 grid((L.M + BM - 1) / BM, (L.g.N + BN - 1) / BN);
    gemm_fp32_kernel<<<grid, NT, 0, stream>>>(L);
    return 1;
}

}  // namespace pnn
