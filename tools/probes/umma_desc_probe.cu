// Hardware probe (developer aid, not part of the library): how does tcgen05.mma address a K-major SWIZZLE_128B
// operand whose descriptor start is NOT 1024-byte aligned and whose stride between 8-row groups (SBO) is not a
// multiple of 1024?  This decides whether a convolution can keep ONE halo tile in shared memory and reach its taps
// by moving the descriptor start (tap reuse) instead of re-fetching a shifted tile per tap.
//
// Shared memory holds 512 logical rows (pixels) of 64 bf16, row i at base + 128 i, its 16-byte chunks XOR-swizzled by
// ((address >> 7) & 7) -- the image TMA writes.  B = 64 x 64 identity, so D[m][n] = A[row(m)][n]: the accumulator shows
// which row and which k the tensor core read for every m.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -o umma_desc_probe umma_desc_probe.cu ; run on a B200.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Probe {
    int row_offset;    // descriptor start = base + 128 * row_offset
    int sbo_bytes;     // stride between 8-row groups
    int base_offset;   // descriptor field (bits 49-51)
    int mode;          // 0: value = row id (i % 256), 1: value = k + 64 * (i & 1)
};

__global__ void __launch_bounds__(128) probe_kernel(Probe P, float* out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sm = raw + (base - smem_u32(raw));
    uint8_t* sa = sm;                  // 512 rows x 128 B
    uint8_t* sb = sm + 512 * 128;      // 64 rows x 128 B
    for (int idx = threadIdx.x; idx < 512 * 64; idx += 128) {
        const int i = idx >> 6, j = idx & 63;
        const float v = P.mode == 0 ? (float)(i % 256) : (float)(j + 64 * (i & 1));
        const uint32_t row_addr = base + i * 128;
        const int chunk = (j >> 3) ^ ((row_addr >> 7) & 7);
        *(__nv_bfloat16*)(sa + i * 128 + chunk * 16 + (j & 7) * 2) = __float2bfloat16_rn(v);
    }
    for (int idx = threadIdx.x; idx < 64 * 64; idx += 128) {
        const int n = idx >> 6, k = idx & 63;
        const int chunk = (k >> 3) ^ (n & 7);
        *(__nv_bfloat16*)(sb + n * 128 + chunk * 16 + (k & 7) * 2) = __float2bfloat16_rn(n == k ? 1.f : 0.f);
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_hi = ((uint32_t)P.sbo_bytes >> 4) | (1u << 14) | ((uint32_t)P.base_offset << 17) | (2u << 29);
        const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t a_start = base + 128 * P.row_offset;
        for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t a_lo = (((a_start + 32 * k4) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lo = (((base + 512 * 128 + 32 * k4) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t accumulate = k4 > 0;
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                ".reg .b64 da, db;\n\t"
                "mov.b64 da, {%1, %2};\n\t"
                "mov.b64 db, {%3, %4};\n\t"
                "setp.ne.b32 p, %6, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
                "}" ::"r"(tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "W_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra W_DONE;\n\t"
        "bra W_LOOP;\n\t"
        "W_DONE:\n\t"
        "}" ::"r"(smem_u32(&bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + half * 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + half * 32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
    }
}

int main() {
    const int smem = 512 * 128 + 64 * 128 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float* d_out;
    cudaMalloc(&d_out, 128 * 64 * sizeof(float));
    std::vector<float> rows(128 * 64), ks(128 * 64);
    const int offsets[] = {0, 1, 2, 3, 8, 11, 21};
    const int sbos[] = {1024, 1152, 1280, 2048, 2304};
    for (int sbo : sbos) {
        for (int off : offsets) {
            for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
                const int bo = bo_mode ? (off & 7) : 0;
                if (bo_mode && bo == 0) continue;
                for (int mode = 0; mode < 2; ++mode) {
                    Probe P{off, sbo, bo, mode};
                    probe_kernel<<<1, 128, smem>>>(P, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) {
                        printf("CUDA error %s (off %d sbo %d bo %d)\n", cudaGetErrorString(e), off, sbo, bo);
                        return 1;
                    }
                    cudaMemcpy((mode ? ks : rows).data(), d_out, 128 * 64 * sizeof(float), cudaMemcpyDeviceToHost);
                }
                // hypothesis: row(m) = off + (m / 8) * (sbo / 128) + m % 8, k = n (absolute-address swizzle)
                int bad_row = 0, bad_k = 0, mixed = 0;
                for (int m = 0; m < 128; ++m) {
                    const int want = off + (m / 8) * (sbo / 128) + (m % 8);
                    for (int n = 0; n < 64; ++n) {
                        if (rows[m * 64 + n] != rows[m * 64]) ++mixed;
                        if ((int)rows[m * 64 + n] != want % 256) ++bad_row;
                        if ((int)ks[m * 64 + n] != n + 64 * (want & 1)) ++bad_k;
                    }
                }
                printf("sbo %4d off %2d base_offset %d : %s (bad_row %d bad_k %d mixed %d)", sbo, off, bo,
                       (bad_row == 0 && bad_k == 0) ? "ABSOLUTE-ADDRESS MODEL OK" : "differs", bad_row, bad_k, mixed);
                if (bad_row || bad_k) {
                    printf("  rows m=0..11:");
                    for (int m = 0; m < 12; ++m) printf(" %d", (int)rows[m * 64]);
                    printf(" | k of m=0,1 (n=0,8,16,24):");
                    for (int m = 0; m < 2; ++m)
                        for (int n = 0; n < 32; n += 8) printf(" %d", (int)ks[m * 64 + n]);
                }
                printf("\n");
            }
        }
    }
    return 0;
}
