// In-loop (batch-1) fully-connected nets, widths 4 and 8 (reference TComPrediction.cpp(substitution):556-614 runs the FC
// graphs with a batch of one, 94 % of the codec's PNN calls).
//
// Arithmetic: fp32 FMA, weights exactly as stored.  The 1200 hidden units are cut over FCI_NCTA = 148 CTAs (16 CTAs own 9
// columns, 132 own 8).  Inside a CTA thread t accumulates rows k = t, t + 256, ... of all the CTA's columns (k ascending),
// the 256 partial sums of a column are added by a fixed shuffle tree per warp and then warp 0..7 in order.  The order
// depends on nothing but K, so the two ways of running a net give the SAME BITS:
//   * fci_persist_kernel: persistent, the weights of BOTH nets resident in shared memory (94 KB + 103 KB per CTA, loaded
//     once), requests arrive through a doorbell in mapped pinned host memory, results leave the same way;
//   * fci_layer_kernel / fci_output_kernel: one launch per layer (replayed as a CUDA graph), weights read from the same
//     per-CTA images in global memory -- the fall-back when 148 CTAs cannot be co-resident, and what the multi-call API
//     uses next to the convolutional graphs.
// Encoder and decoder reconstructions therefore match bit for bit whichever of the two serves a call.
#include "kernels_common.cuh"

#include <cstdlib>
#include <stdexcept>

namespace pnn {

__device__ __forceinline__ int fci_cols(int cta) { return 8 + (cta < FCI_NCTA_WIDE ? 1 : 0); }
__device__ __forceinline__ int fci_col0(int cta) { return cta * 8 + (cta < FCI_NCTA_WIDE ? cta : FCI_NCTA_WIDE); }

// ---------------------------------------------------------------------------------------------
// {payload, tag} pairs: every value travels with the tag of the call (and layer) it belongs to in ONE 8-byte store
// (single-copy atomic), the consumer polls the pair until the tag matches.  No fence, no counter, no second round trip.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store_gpu(uint2* p, unsigned bits, unsigned tag) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(bits), "r"(tag) : "memory");
}
__device__ __forceinline__ void ll_store_sys(uint2* p, unsigned bits, unsigned tag) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(bits), "r"(tag) : "memory");
}
template <bool SYS>
__device__ __forceinline__ uint2 ll_load(const uint2* p) {
    uint2 v;
    if (SYS) asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------------
// Loads from a weight image that lives either in shared memory (persistent kernel) or in global memory (layer kernels).
// Explicit state spaces: a generic pointer would cost LD.E (~100 cycles) instead of LDS (~30).
// ---------------------------------------------------------------------------------------------
template <bool SMEM>
__device__ __forceinline__ float4 img_ld4(const float* p) {
    float4 v;
    if (SMEM) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    } else {
        asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    }
    return v;
}
template <bool SMEM>
__device__ __forceinline__ float img_ld1(const float* p) {
    float v;
    if (SMEM) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    else asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// Packed image of one CTA (floats), K0 = inputs of the first layer:
//   main0 [2][K0][4] | main1 [2][1200][4] | main2 [2][1200][4] | extra0 [K0] | extra1 [1200] | extra2 [1200] | out [1200] | bias [3][9] + [1]
// main = the CTA's first 8 columns as two planes of 4 (a warp's 16-byte loads of consecutive rows are then contiguous:
// no shared-memory bank conflict), extra = the 9th column of the 16 wide CTAs.
__device__ __forceinline__ int fci_rows_before(int K0, int layer) { return layer == 0 ? 0 : (layer == 1 ? K0 : K0 + FCI_HID); }
__device__ __forceinline__ const float* fci_main_ptr(const float* img, int K0, int layer) { return img + 8 * fci_rows_before(K0, layer); }
__device__ __forceinline__ const float* fci_extra_ptr(const float* img, int K0, int layer) {
    return img + 8 * (K0 + 2 * FCI_HID) + fci_rows_before(K0, layer);
}
__device__ __forceinline__ const float* fci_out_ptr(const float* img, int K0) { return img + 9 * (K0 + 2 * FCI_HID); }
__device__ __forceinline__ const float* fci_bias_ptr(const float* img, int K0, int layer) {
    return img + 9 * (K0 + 2 * FCI_HID) + FCI_HID + layer * FCI_COLS_MAX;
}

// ---------------------------------------------------------------------------------------------
// One hidden layer for the `cols` columns of a CTA.  Thread t owns rows k = t + 256 j (j < 5) of the CTA's weight slice:
// it loads them into registers (FciRows) BEFORE its inputs x[k] are there -- in the persistent kernel the inputs are the
// very {value, tag} pairs thread t polls, so activations never pass through shared memory and the weight loads hide behind
// the wait.  fci_hidden_sum leaves the per-warp sums in `wred`; after a __syncthreads any thread may finish column c with
// fci_hidden_finish (warps 0..7 in order, bias, LeakyReLU).  Fixed order throughout.
// ---------------------------------------------------------------------------------------------
struct FciRows {
    float4 w0[5], w1[5];
    float we[5];
};

template <bool SMEM>
__device__ __forceinline__ void fci_load_rows(FciRows& r, const float* main, const float* extra, int K, int cols, int t) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int k = t + 256 * j;
        r.w0[j] = r.w1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        r.we[j] = 0.f;
        if (k < K) {
            r.w0[j] = img_ld4<SMEM>(main + 4 * k);
            r.w1[j] = img_ld4<SMEM>(main + 4 * (K + k));
            if (cols > 8) r.we[j] = img_ld1<SMEM>(extra + k);
        }
    }
}

__device__ __forceinline__ void fci_hidden_sum(const FciRows& r, const float (&x)[5], int K, int cols, float (*wred)[FCI_COLS_MAX], int t) {
    float acc[FCI_COLS_MAX];
#pragma unroll
    for (int c = 0; c < FCI_COLS_MAX; ++c) acc[c] = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        if (t + 256 * j < K) {
            acc[0] = fmaf(x[j], r.w0[j].x, acc[0]);
            acc[1] = fmaf(x[j], r.w0[j].y, acc[1]);
            acc[2] = fmaf(x[j], r.w0[j].z, acc[2]);
            acc[3] = fmaf(x[j], r.w0[j].w, acc[3]);
            acc[4] = fmaf(x[j], r.w1[j].x, acc[4]);
            acc[5] = fmaf(x[j], r.w1[j].y, acc[5]);
            acc[6] = fmaf(x[j], r.w1[j].z, acc[6]);
            acc[7] = fmaf(x[j], r.w1[j].w, acc[7]);
            acc[8] = fmaf(x[j], r.we[j], acc[8]);
        }
    }
    // Sum over the 32 lanes, 8 columns at once: at offsets 16, 8, 4 a lane keeps one half of its columns and hands the
    // other half to its partner (4 + 2 + 1 exchanges), offsets 2 and 1 finish the one column that is left: 9 shuffles
    // instead of 40.  Lanes 0, 4, ..., 28 then hold the warp's sum of column ((l >> 4) & 1) * 4 + ((l >> 3) & 1) * 2 +
    // ((l >> 2) & 1).
    const int lane = t & 31;
    float h4[4], h2[2], h1;
    {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float keep = up ? acc[4 + i] : acc[i], give = up ? acc[i] : acc[4 + i];
            h4[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
        }
    }
    {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float keep = up ? h4[2 + i] : h4[i], give = up ? h4[i] : h4[2 + i];
            h2[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
        }
    }
    {
        const bool up = (lane & 4) != 0;
        const float keep = up ? h2[1] : h2[0], give = up ? h2[0] : h2[1];
        h1 = keep + __shfl_xor_sync(0xffffffffu, give, 4);
    }
    h1 += __shfl_xor_sync(0xffffffffu, h1, 2);
    h1 += __shfl_xor_sync(0xffffffffu, h1, 1);
    if (cols > 8) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[8] += __shfl_xor_sync(0xffffffffu, acc[8], o);
    }
    if ((lane & 3) == 0) wred[t >> 5][((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = h1;
    if (lane == 0) wred[t >> 5][8] = acc[8];
}

__device__ __forceinline__ float fci_hidden_finish(const float (*wred)[FCI_COLS_MAX], const float* bias, int c) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += wred[w][c];
    return leaky_relu(s + bias[c]);
}

// Output layer: one CTA per output, thread t owns weights k = t + 256 j.  The sum is returned in thread 0 (bias not added).
template <bool SMEM>
__device__ __forceinline__ void fci_load_out(float (&w)[5], const float* w3, int t) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int k = t + 256 * j;
        w[j] = k < FCI_HID ? img_ld1<SMEM>(w3 + k) : 0.f;
    }
}
__device__ __forceinline__ float fci_output_sum(const float (&w)[5], const float (&x)[5], float* wsum, int t) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        if (t + 256 * j < FCI_HID) a = fmaf(w[j], x[j], a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((t & 31) == 0) wsum[t >> 5] = a;
    __syncthreads();
    float s = 0.f;
    if (t == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s += wsum[i];
    }
    return s;
}

// x[j] <- payload of pair t + 256 j (< K) once its tag is `tag`; a thread asks again only for the pairs it still misses
template <bool SYS>
__device__ __forceinline__ void ll_wait(const uint2* pairs, int K, unsigned tag, float (&x)[5], int t) {
    unsigned pending = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        pending |= (t + 256 * j < K) ? 1u << j : 0u;
        x[j] = 0.f;
    }
    while (pending) {
        uint2 v[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            if (pending & (1u << j)) v[j] = ll_load<SYS>(pairs + t + 256 * j);
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            if ((pending & (1u << j)) && v[j].y == tag) {
                x[j] = __uint_as_float(v[j].x);
                pending &= ~(1u << j);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent kernel.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int fci_smem_floats(int img0, int img1) { return img0 + img1 + 8 * FCI_COLS_MAX + 8 + FCI_CTX_MAX + 8; }

__device__ __forceinline__ void fci_stamp(const FciPersist& P, int i) {
    // SM cycle counter (cheap); slot 9 / 10: %globaltimer at the first / last stamp, to turn cycles into nanoseconds
    if (P.stamps && threadIdx.x == 0) {
        P.stamps[i] = (unsigned long long)clock64();
        if (i == 0 || i == 8) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            P.stamps[i == 0 ? 9 : 10] = t;
        }
    }
}

// Number of request payloads after the header: 10-bit codes travel three per payload, raw floats one per payload.
__device__ __forceinline__ int fci_payloads(unsigned cmd, int K0) { return (cmd & FCI_CMD_FLOAT) ? K0 : (K0 + 2) / 3; }

__global__ void __launch_bounds__(256, 1) fci_persist_kernel(const __grid_constant__ FciPersist P) {
    extern __shared__ __align__(16) float fci_smem[];
    const int t = threadIdx.x, cta = blockIdx.x;
    const int cols = fci_cols(cta), col0 = fci_col0(cta);
    const int stride0 = P.net[0].present ? P.net[0].stride : 0, stride1 = P.net[1].present ? P.net[1].stride : 0;
    float(*wred)[FCI_COLS_MAX] = reinterpret_cast<float(*)[FCI_COLS_MAX]>(fci_smem + stride0 + stride1);
    float* wsum = fci_smem + stride0 + stride1 + 8 * FCI_COLS_MAX;
    uint32_t* pay_s = reinterpret_cast<uint32_t*>(wsum + 8);           // [FCI_CTX_MAX] request payloads (gate)
    uint32_t* cmd_s = pay_s + FCI_CTX_MAX;
    // the weights of this CTA's columns, both nets, once
    for (int n = 0; n < 2; ++n) {
        if (!P.net[n].present) continue;
        const float4* src = reinterpret_cast<const float4*>(P.net[n].images + (size_t)cta * P.net[n].stride);
        float4* dst = reinterpret_cast<float4*>(fci_smem + (n ? stride0 : 0));
        for (int i = t; i < P.net[n].stride / 4; i += 256) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    // every consumer reads one of FCI_COPIES identical copies of what it waits for: 148 CTAs polling the same few KB would
    // queue on the handful of L2 slices that hold them
    const int copy = cta & (FCI_COPIES - 1);
    const uint2* relay_in = P.relay + (size_t)copy * FCI_REQ_PAIRS;
    // the thread that finishes column c of a layer for copy r of the exchange buffer
    const int fin_r = t / cols, fin_c = t - fin_r * cols;
    const bool fin = t < cols * FCI_COPIES;
    unsigned seq = P.seq0;
    for (;; seq = (seq + 1u) & 0x3fffffffu, seq = seq ? seq : 1u) {
        unsigned cmd;
        float x[5];
        FciRows rows;
        if (cta == 0) {
            // ---- gate: the request arrives in mapped host memory.  Warp 0 polls its first 32 pairs (header + 31 payloads:
            // a whole 4x4 context) with one 256-byte read per round trip; longer requests are completed by all warps.
            if (t < 32) {
                for (;;) {
                    const uint2 v = ll_load<true>(P.req + t);
                    const unsigned c = __shfl_sync(0xffffffffu, v.x, 0);
                    bool ok = v.y == seq;
                    if (__shfl_sync(0xffffffffu, (int)ok, 0)) {                    // header of this request seen
                        const int need = (c & FCI_CMD_QUIT) ? 0 : fci_payloads(c, P.net[c & 1].K0);
                        ok = ok || t > need;                                         // lanes beyond the request do not wait
                        if (__all_sync(0xffffffffu, ok)) {
                            if (t == 0) cmd_s[0] = c;
                            else if (t <= need) pay_s[t - 1] = v.x;
                            break;
                        }
                    }
                }
            }
            __syncthreads();
            cmd = cmd_s[0];
            if (!(cmd & FCI_CMD_QUIT)) {
                const int need = fci_payloads(cmd, P.net[cmd & 1].K0);
                if (need > 31) {
                    // pairs 32 .. need of the request (written before the header): one more round trip
                    for (int i = 32 + t; i <= need; i += 256) {
                        uint2 v;
                        do { v = ll_load<true>(P.req + i); } while (v.y != seq);
                        pay_s[i - 1] = v.x;
                    }
                    __syncthreads();
                }
            }
            fci_stamp(P, 0);
            if (t < FCI_COPIES) ll_store_gpu(P.relay + (size_t)t * FCI_REQ_PAIRS, cmd, seq);
            if (cmd & FCI_CMD_QUIT) break;
            // decode (mean subtraction of the codec's context, reference extraction_context.cpp:56-90; masked pixels
            // arrive as code 0x100 -> 0) and hand the context to the other CTAs
            const int K0 = P.net[cmd & 1].K0;
#pragma unroll
            for (int j = 0; j < 5; ++j) x[j] = 0.f;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int e = t + 256 * j;
                if (e < K0) {
                    float v;
                    if (cmd & FCI_CMD_FLOAT) {
                        v = __uint_as_float(pay_s[e]);
                    } else {
                        const unsigned code = (pay_s[e / 3] >> (10 * (e % 3))) & 0x3ffu;
                        v = code >= 0x100u ? 0.f : (float)code - P.mean;
                    }
                    x[j] = v;
#pragma unroll
                    for (int r = 0; r < FCI_COPIES; ++r) ll_store_gpu(P.relay + (size_t)r * FCI_REQ_PAIRS + 1 + e, __float_as_uint(v), seq);
                }
            }
            const float* img = fci_smem + ((cmd & 1) ? stride0 : 0);
            fci_load_rows<true>(rows, fci_main_ptr(img, K0, 0), fci_extra_ptr(img, K0, 0), K0, cols, t);
        } else {
            // ---- the others: one lane polls the header of its copy, then every thread waits for its rows of the context
            if (t == 0) {
                uint2 v;
                do { v = ll_load<false>(relay_in); } while (v.y != seq);
                cmd_s[0] = v.x;
            }
            __syncthreads();
            cmd = cmd_s[0];
            if (cta == 1) fci_stamp(P, 1);
            if (cmd & FCI_CMD_QUIT) break;
            const int K0 = P.net[cmd & 1].K0;
            const float* img = fci_smem + ((cmd & 1) ? stride0 : 0);
            fci_load_rows<true>(rows, fci_main_ptr(img, K0, 0), fci_extra_ptr(img, K0, 0), K0, cols, t);
            ll_wait<false>(relay_in + 1, K0, seq, x, t);
        }
        const int n = cmd & 1;
        const int K0 = P.net[n].K0, N3 = P.net[n].N3;
        const float* img = fci_smem + (n ? stride0 : 0);
        const unsigned tag0 = seq << 2;
        bool has_output = true;
        for (int layer = 0; layer < 3; ++layer) {
            const int K = layer == 0 ? K0 : FCI_HID;
            fci_hidden_sum(rows, x, K, cols, wred, t);
            __syncthreads();
            const unsigned tag = tag0 + (unsigned)layer + 1u;
            if (fin) {
                const float y = fci_hidden_finish(wred, fci_bias_ptr(img, K0, layer), fin_c);
                ll_store_gpu(P.xchg + ((size_t)layer * FCI_COPIES + fin_r) * FCI_HID + col0 + fin_c, __float_as_uint(y), tag);
            }
            if (cta == 0) fci_stamp(P, 2 + 2 * layer);
            if (layer == 2 && cta >= N3) {                   // only the CTAs that own an output run the last layer
                has_output = false;
                break;
            }
            // the weights of the next layer travel to registers while its inputs are on their way
            float w3[5];
            if (layer < 2) fci_load_rows<true>(rows, fci_main_ptr(img, K0, layer + 1), fci_extra_ptr(img, K0, layer + 1), FCI_HID, cols, t);
            else fci_load_out<true>(w3, fci_out_ptr(img, K0), t);
            ll_wait<false>(P.xchg + ((size_t)layer * FCI_COPIES + copy) * FCI_HID, FCI_HID, tag, x, t);
            if (cta == 0) fci_stamp(P, 3 + 2 * layer);
            if (layer == 2) {
                const float s = fci_output_sum(w3, x, wsum, t);
                if (t == 0) {
                    const float p = s + fci_bias_ptr(img, K0, 3)[0];
                    const float v = fminf(fmaxf(p + P.mean, 0.f), 255.f);
                    const float r = P.round_mode == 0 ? rintf(v) : roundf(v);
                    ll_store_sys(P.out_ll + cta, __float_as_uint(p), seq);
                    ll_store_sys(P.out_ll + 64 + cta, (unsigned)(int)r, seq);
                }
                if (cta == 0) fci_stamp(P, 8);
            }
            __syncthreads();                                  // wred / wsum are reused by the next layer
        }
        (void)has_output;
    }
}

// ---------------------------------------------------------------------------------------------
// One launch per layer (CUDA-graph fall-back): same device functions, weights from the images in global memory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fci_layer_kernel(FciLayerLaunch L) {
    __shared__ float wred[8][FCI_COLS_MAX];
    const int t = threadIdx.x, cta = blockIdx.x;
    const int cols = fci_cols(cta), col0 = fci_col0(cta);
    const int K = L.layer == 0 ? L.K0 : FCI_HID;
    const float* img = L.images + (size_t)cta * L.stride;
    FciRows rows;
    fci_load_rows<false>(rows, fci_main_ptr(img, L.K0, L.layer), fci_extra_ptr(img, L.K0, L.layer), K, cols, t);
    float x[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int k = t + 256 * j;
        x[j] = k < K ? (L.layer == 0 ? hm_context_value(L.staged, L.W, L.mean, k) : L.x[k]) : 0.f;
    }
    fci_hidden_sum(rows, x, K, cols, wred, t);
    __syncthreads();
    if (t < cols) L.y[col0 + t] = fci_hidden_finish(wred, fci_bias_ptr(img, L.K0, L.layer), t);
}

__global__ void __launch_bounds__(256) fci_output_kernel(FciLayerLaunch L) {
    __shared__ float wsum[8];
    const int t = threadIdx.x, cta = blockIdx.x;                      // grid = N3
    const float* img = L.images + (size_t)cta * L.stride;
    float w3[5], x[5];
    fci_load_out<false>(w3, fci_out_ptr(img, L.K0), t);
#pragma unroll
    for (int j = 0; j < 5; ++j) x[j] = t + 256 * j < FCI_HID ? L.x[t + 256 * j] : 0.f;
    const float s = fci_output_sum(w3, x, wsum, t);
    if (t == 0) final_store(L.fin, cta, s + fci_bias_ptr(img, L.K0, 3)[0]);
}

int launch_fci_layer(const FciLayerLaunch& L, cudaStream_t stream) {
    if (L.layer < 3) fci_layer_kernel<<<FCI_NCTA, 256, 0, stream>>>(L);
    else fci_output_kernel<<<L.N3, 256, 0, stream>>>(L);
    return 1;
}

size_t fci_persist_smem_bytes(const FciPersist& P) {
    return sizeof(float) * (size_t)fci_smem_floats(P.net[0].present ? P.net[0].stride : 0, P.net[1].present ? P.net[1].stride : 0);
}

// Cooperative launch: the runtime refuses it (cudaErrorCooperativeLaunchTooLarge) unless all 148 CTAs can be resident at
// once, which is what the exchange between the layers relies on.
cudaError_t launch_fci_persist(const FciPersist& P, cudaStream_t stream) {
    const size_t smem = fci_persist_smem_bytes(P);
    cudaError_t e = cudaFuncSetAttribute(fci_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    FciPersist copy = P;
    void* args[] = {(void*)&copy};
    return cudaLaunchCooperativeKernel((const void*)fci_persist_kernel, dim3(FCI_NCTA), dim3(256), args, smem, stream);
}

// Per-CTA packed weight images of one FC net (host side; layout above fci_main_ptr).
std::vector<float> fci_build_images(const float* w0, const float* w1, const float* w2, const float* w3, const float* b0,
                                    const float* b1, const float* b2, const float* b3, int K0, int N3) {
    if (N3 > 64 || N3 > FCI_NCTA || K0 > FCI_CTX_MAX) throw std::runtime_error("the in-loop FC kernels serve widths 4 and 8 only");
    const int stride = fci_image_floats(K0);
    std::vector<float> img((size_t)FCI_NCTA * stride, 0.f);
    const float* ws[3] = {w0, w1, w2};
    const float* bs[3] = {b0, b1, b2};
    for (int cta = 0; cta < FCI_NCTA; ++cta) {
        const int cols = 8 + (cta < FCI_NCTA_WIDE ? 1 : 0), col0 = cta * 8 + (cta < FCI_NCTA_WIDE ? cta : FCI_NCTA_WIDE);
        float* base = img.data() + (size_t)cta * stride;
        float* main = base;
        float* extra = base + 8 * (size_t)(K0 + 2 * FCI_HID);
        for (int layer = 0; layer < 3; ++layer) {
            const int K = layer == 0 ? K0 : FCI_HID;
            for (int k = 0; k < K; ++k) {
                for (int c = 0; c < 8; ++c) main[(size_t)(c / 4) * K * 4 + (size_t)k * 4 + c % 4] = ws[layer][(size_t)k * FCI_HID + col0 + c];
                if (cols > 8) extra[k] = ws[layer][(size_t)k * FCI_HID + col0 + 8];
            }
            main += (size_t)K * 8;
            extra += K;
        }
        float* out = base + 9 * (size_t)(K0 + 2 * FCI_HID);
        if (cta < N3) {
            for (int k = 0; k < FCI_HID; ++k) out[k] = w3[(size_t)k * N3 + cta];
        }
        float* bias = out + FCI_HID;
        for (int layer = 0; layer < 3; ++layer)
            for (int c = 0; c < cols; ++c) bias[layer * FCI_COLS_MAX + c] = bs[layer][col0 + c];
        if (cta < N3) bias[3 * FCI_COLS_MAX] = b3[cta];
    }
    return img;
}

}  // namespace pnn
