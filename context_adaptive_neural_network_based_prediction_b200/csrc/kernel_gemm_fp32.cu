// fp32 FFMA implicit-GEMM kernel: the accuracy yard-stick of the GEMM-shaped layers (FC, convolutions,
// transposed-convolution phases).  Same geometry and reduction structure (k ascending) as the tcgen05
// kernel, plain fp32 arithmetic.  Not the throughput path.
#include "kernels_common.cuh"

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

}  // namespace

int launch_gemm_fp32(const GemmLaunch& L, cudaStream_t stream) {
    if (L.M == 0) return 0;
    dim3 grid((L.M + BM - 1) / BM, (L.g.N + BN - 1) / BN);
    gemm_fp32_kernel<<<grid, NT, 0, stream>>>(L);
    return 1;
}

}  // namespace pnn
