// Internal declarations shared by the host side and the kernels of libpnn_cuda.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

namespace pnn {

// ---------------------------------------------------------------------------------------------
// Activation storage.  fp32 precision: p0 = float*.  bf16x3 precision: p0 = hi plane, p1 = lo plane
// (__nv_bfloat16*), value = hi + lo.  Layout is always [sample][pixel][channel] (NHWC), exact pitches.
// ---------------------------------------------------------------------------------------------
struct Act {
    void* p0;
    void* p1;
};

// Output conversion of a layer.
enum OutMode {
    OUT_ACT = 0,    // next layer's input (fp32 or split planes depending on precision)
    OUT_FINAL = 1   // network output: raw fp32 (+ optional rounded u8 / int32), see FinalOut
};

struct FinalOut {
    float* raw;        // [n, W*W] prediction before the mean is added back (may be null)
    uint8_t* u8;       // [n, W*W] round(clip(p + mean, 0, 255)) (may be null)
    int32_t* i32;      // same as int32 (HM's Pel) (may be null)
    float mean;
    int round_mode;    // PNN_ROUND_*
};

// Geometry of one GEMM-shaped layer (FC, convolution, or one phase of a transposed convolution):
//   out[b, oy*osy+ooy, ox*osx+oox, n] = act( sum_{ty,tx,ci} in[b, oy*sy_o+ty*sy_t+cy, ox*sx_o+tx*sx_t+cx, ci]
//                                                         * Wm[(ty*TW+tx)*Cin+ci, n] + bias[n] )
// rows m = b*P + oy*OW + ox.  Out-of-range input pixels read as zero (SAME padding / tconv borders).
struct GemmGeom {
    int P, OW;                  // output positions per sample in this launch, and their row width
    int Cin, TH, TW;            // K = TH*TW*Cin
    int IH, IW;                 // input map
    int sy_o, sy_t, cy;
    int sx_o, sx_t, cx;
    int64_t in_sample_stride;   // elements
    int N, K;
    int OHf, OWf;               // full output map (for the output address)
    int osy, ooy, osx, oox;
    int64_t out_sample_stride;  // elements
    int leaky;                  // 1: LeakyReLU(0.1), 0: linear
};

struct GemmLaunch {
    GemmGeom g;
    Act in;
    Act out;                    // OUT_ACT
    FinalOut fin;               // OUT_FINAL
    int out_mode;
    int M;                      // rows in this launch
    const float* w_fp32;        // [K][N] fp32 (fp32 precision)
    const uint8_t* w_tiles;     // pre-swizzled bf16 hi/lo tiles (bf16x3 precision)
    const float* bias;          // [N]
    int debug_flags;            // tuning aid (pnn_debug_time_gemm): bit 0 skip A loads, 1 skip B loads, 2 skip epilogue stores,
                                // 3 no proxy fence, 4 one K step per block, 5 plain arrive instead of tcgen05.commit, 6 issuer skips waits
    // split-K (in-loop batch-1 path only, where 1-2 tiles would otherwise walk K serially): slice ks of the K blocks
    // writes its raw fp32 accumulators to partial[ks][M][N]; splitk_reduce adds the slices in ascending order, then
    // bias / LeakyReLU / hi-lo split.  1 = off.
    int split_k;
    float* partial;
    // TMA path for the A operand of real convolutions (Cin % 64 == 0, more than one tap): the M tile is a box of
    // bw x bh output pixels x nb samples (bw * bh * nb = 128, powers of two) loaded by ONE cp.async.bulk.tensor per
    // plane and K block (hardware 128B swizzle, zero fill for SAME padding, element strides for stride 2).
    int tma;                    // 1: tile = (sample tile, y tile, x tile, n tile), rows in box order
    int bw_log2, bh_log2, x_tiles, y_tiles;
    // Tap reuse along x (TMA path, bw = 8, N <= 128): the taps of one kernel row whose input columns differ by whole
    // box elements share ONE box that is 2 pixels wider (10 x bh x nb rows of 128 B); tap t is reached by starting
    // the tcgen05 shared-memory descriptor xr_off rows further (the 128B swizzle is a function of the absolute
    // shared-memory address, tools/probes/umma_desc_probe.cu), with 10 rows between the 8-row groups.  K order:
    // (ty, 64-channel chunk, group, tap).  Stride-2 convolutions have two groups (even / odd columns).
    int xr;                     // 1: on
    int xr_groups;
    int xr_start[2];            // first input column of the group's box relative to ox0 * sx_o
    int xr_ntaps[2];
    int xr_tx[2][3];
    int xr_off[2][3];
};
int launch_splitk_reduce(const GemmLaunch& L, cudaStream_t stream);

// bf16x3 weight tiling (see DESIGN.md "weight tiles"): N is cut in tiles of TC_BN rows (the last one
// narrower, a multiple of 16), K in blocks of 64.  Tile (nt, kb) is stored as the exact shared-memory
// image of its hi plane followed by its lo plane, 128-byte-swizzled K-major, so that one bulk copy
// lands it ready for tcgen05.mma.
constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_BK = 64;

inline int tc_num_kb(int K) { return (K + TC_BK - 1) / TC_BK; }
inline int tc_num_nt(int N) { return (N + TC_BN - 1) / TC_BN; }
inline int tc_tile_bn(int N, int nt) {
    int rem = N - nt * TC_BN;
    if (rem > TC_BN) rem = TC_BN;
    return (rem + 15) / 16 * 16;
}
// byte offset of tile (nt, kb)
inline size_t tc_tile_offset(int N, int K, int nt, int kb) {
    size_t off = 0;
    for (int t = 0; t < nt; ++t) off += (size_t)tc_num_kb(K) * 2 * tc_tile_bn(N, t) * 128;
    return off + (size_t)kb * 2 * tc_tile_bn(N, nt) * 128;
}
inline size_t tc_total_bytes(int N, int K) { return tc_tile_offset(N, K, tc_num_nt(N), 0); }

// [K][N] fp32 weights (device) -> the tile image described above (device); kernels_small.cu
int launch_make_tc_tiles(const float* w_kn, int K, int N, uint8_t* tiles, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// Kernel launchers (implemented in the .cu files).  All asynchronous on `stream`.
// Each returns the number of kernels it launched.
// ---------------------------------------------------------------------------------------------
int launch_gemm_fp32(const GemmLaunch& L, cudaStream_t stream);
int launch_gemm_skinny(const GemmLaunch& L, cudaStream_t stream);   // batch-1 (in-loop) fp32 layer: 16 x 16 output tiles over the whole K
int launch_gemm_tc(const GemmLaunch& L, cudaStream_t stream);   // fills the TMA fields itself when the path applies
void gemm_tc_set_tma(int enabled);
// one-time opt-in for the tcgen05 kernel's dynamic shared memory
cudaError_t gemm_tc_init();

// First convolution of a branch (one input channel) as a direct fp32 FFMA kernel fused with bias, LeakyReLU and the
// hi/lo split: replaces im2col + GEMM (K = k*k is too small for the tensor pipe to matter; the layer is bound by
// writing its C-channel output).
struct ConvFirstWeights {
    float w[25 * 64];    // [tap][64]: TF layout [k, k, 1, C], rows padded to 64 channels
    float b[64];
};
struct ConvFirstLaunch {
    const float* in;     // [n, IH, IW] fp32
    Act out;             // [n*OH*OW, C]
    int n, IH, IW, OH, OW, C, k, stride, pad, split;
    int in_loop;         // 1: batch-1 call of the codec (keeps the direct fp32 FFMA kernel, like PNN_PRECISION_FP32)
};
// Batched bf16x3 calls: one tensor-core launch for all channels.  fp32 precision and in-loop calls: direct FFMA kernels,
// one launch per group of 16 output channels (or one in all for small grids).  Returns the number of launches.
int launch_conv_first(const ConvFirstLaunch& L, const ConvFirstWeights& W, cudaStream_t stream);

// col2im of the last transposed convolution + the output epilogue: D[(b, iy, ix), ky*k+kx] holds the
// per-tap dot products (GEMM output); out[y, x] = bias + sum over the taps with iy*s + ky - pad == y.
struct Col2imLaunch {
    Act d;               // [n*IH*IW, NP]
    float bias;
    FinalOut fin;
    int n, IH, IW, k, stride, pad, NP, split;
};
int launch_col2im(const Col2imLaunch& L, cudaStream_t stream);

// Channel-wise fully-connected merger (reference pnn/tfutils.py:8-73) + LeakyReLU.
struct MergerLaunch {
    Act in0, in1;        // [n, 48, C], [n, 32, C]
    Act out;             // [n, 16, C]
    const float* w;      // transposed on the host to [80][16][C]
    const float* bias;   // [16][C]
    int n, C, split;
    int in_loop;         // 1: batch-1 call of the codec (fp32 direct kernel; the batched paths keep one arithmetic for every n)
};
int launch_merger(const MergerLaunch& L, cudaStream_t stream);

// Fused gather: uint8 image -> mean-centred, masked context (reference sets/common.py:99-109, 454-472).
struct GatherLaunch {
    const uint8_t* images;      // [n_images, H, Wimg]
    const int32_t* image_index; // may be null
    const int32_t* rows;
    const int32_t* cols;
    int64_t n;
    int H, Wimg, W, mask_w, mask_h, n_images;
    float mean;
    // destination of the above portion [n][W*3W] and of the left portion [n][2W*W]; `pitch_*` in elements
    Act above, left;
    int64_t pitch_above, pitch_left;
    int split;
};
int launch_gather_image(const GatherLaunch& L, cudaStream_t stream);

// HM gather: staged int32 context -> masked, mean-centred (reference extraction_context.cpp:56-205).
// The per-call availability information travels in a 4-int header in front of the pixels, so that kernel
// parameters are identical for every call and the launch sequence can be replayed as a CUDA graph:
//   staged[0], staged[1]  bit i: above / above-right unit i available (units of staged[2] columns)
//   staged[2]             unit width; 0 = the pixels are float bits of an already pre-processed context
//   staged[3]             rows of the left portion that are copied (the rest stays zero)
//   staged[4 ...]         [W*3W] above rows then [2W*W] left rows, raw reconstruction pixels
constexpr int HM_HEADER_INTS = 4;
struct GatherHmLaunch {
    const int32_t* staged;
    int W;
    float mean;
    Act above, left;
    int split;
};
int launch_gather_hm(const GatherHmLaunch& L, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// In-loop (batch-1) fully-connected nets, widths 4 and 8 (kernel_fc_inloop.cu).  The 1200 hidden units are cut over
// FCI_NCTA CTAs (the first FCI_NCTA_WIDE own 9 columns, the others 8); every CTA has a packed image of its weights:
//   L0 [K0][cols] | L1 [1200][cols] | L2 [1200][cols] | L3 [1200] (CTA o < N3 owns output o) | bias [3][cols] + [1]
// at a fixed stride of fci_image_floats(K0) floats.
// ---------------------------------------------------------------------------------------------
constexpr int FCI_NCTA = 148;
constexpr int FCI_NCTA_WIDE = 16;        // 16 * 9 + 132 * 8 = 1200
constexpr int FCI_HID = 1200;
constexpr int FCI_COLS_MAX = 9;
constexpr int FCI_CTX_MAX = 320;         // 5 * 8 * 8
constexpr int FCI_REQ_PAIRS = 352;       // header pair + up to 320 payload pairs, padded to a multiple of 256 bytes
constexpr int FCI_COPIES = 8;            // identical copies of everything 148 CTAs poll at once (spreads the L2 slices)
inline int fci_image_floats(int K0) { return (K0 + 2 * FCI_HID) * FCI_COLS_MAX + FCI_HID + 32; }
std::vector<float> fci_build_images(const float* w0, const float* w1, const float* w2, const float* w3, const float* b0,
                                    const float* b1, const float* b2, const float* b3, int K0, int N3);

// Request header (payload of pair 0):
constexpr unsigned FCI_CMD_NET = 1u;     // bit 0: 0 = the width-4 net, 1 = the width-8 net
constexpr unsigned FCI_CMD_FLOAT = 2u;   // payloads are the float bits of a pre-processed context, one per pair; otherwise
                                         // 10-bit codes, three per pair: 0..255 = reconstruction pixel (the kernel subtracts
                                         // the mean), 0x100 = masked / unavailable (-> 0)
constexpr unsigned FCI_CMD_QUIT = 4u;    // the kernel exits

struct FciNet {
    const float* images;    // [FCI_NCTA][stride]
    int K0, N3, W, present, stride;
};
// Persistent kernel: one CTA per SM for the life of the kernel, the images of both nets in shared memory.  A request is a
// header pair + payload pairs {payload, seq} written by the host into mapped pinned memory.  CTA 0 polls the first 32
// pairs (one 256-byte read per PCIe round trip; a whole 4x4 context fits), decodes the context and relays it to the
// other CTAs through device memory; layers hand their 1200 activations over as {bits, tag} pairs; the CTAs that own an
// output write {raw bits, seq} and {rounded int, seq} into mapped pinned memory, which the host polls.
struct FciPersist {
    FciNet net[2];
    const uint2* req;       // mapped host memory [FCI_REQ_PAIRS]
    uint2* relay;           // device memory      [FCI_COPIES][FCI_REQ_PAIRS]
    uint2* xchg;            // device memory      [3][FCI_COPIES][FCI_HID]
    uint2* out_ll;          // mapped host memory [64] raw + [64] rounded
    float mean;
    int round_mode;
    unsigned seq0;          // sequence number of the first request this launch serves
    unsigned long long* stamps;   // tuning aid (PNN_FC_STAMPS=1): %globaltimer at 9 points of a call, or NULL
};
size_t fci_persist_smem_bytes(const FciPersist& P);
cudaError_t launch_fci_persist(const FciPersist& P, cudaStream_t stream);

// One layer per launch (CUDA-graph fall-back, same bits): layer 0 reads the staged context (see GatherHmLaunch),
// layers 1-2 read x[1200], layer 3 (grid = N3) runs the output epilogue.
struct FciLayerLaunch {
    const float* images;
    int K0, N3, W, stride, layer;
    const int32_t* staged;
    const float* x;
    float* y;
    float mean;
    FinalOut fin;
};
int launch_fci_layer(const FciLayerLaunch& L, cudaStream_t stream);
void small_kernels_init();

// Best HEVC intra mode of every block (35 modes on an unfiltered pattern) -- the baseline of the offline evaluation.
int launch_hevc_best_mode(const uint8_t* images, const int32_t* image_index, const int32_t* rows, const int32_t* cols, int64_t n,
                          int H, int Wimg, int W, int mask_w, int mask_h, uint8_t* best_index, double* psnr, uint8_t* pred,
                          cudaStream_t stream, int n_images);

int launch_win_flags(const double* psnr, const double* baseline, int64_t n, uint8_t* win, cudaStream_t stream);

// fp32 -> (fp32 copy | split planes)
int launch_convert_input(const float* src, Act dst, int64_t n_elems, int split, cudaStream_t stream);

// PSNR of each block against its target in the image (reference tools/tools.py:364-401).
int launch_psnr(const uint8_t* images, const int32_t* image_index, const int32_t* rows, const int32_t* cols,
                int64_t n, int H, int Wimg, int W, const uint8_t* pred_u8, double* out, cudaStream_t stream, int n_images);

// ---------------------------------------------------------------- weight files (host side)
// Tensors of one net keyed by the TensorFlow variable names of the reference graph (reference layouts).
struct FlatFile {
    int width = 0;
    bool is_fc = false;
    std::map<std::string, std::vector<float>> t;
    std::map<std::string, std::vector<int>> shape;
};
// Binary GraphDef written by the reference's freezing script (`graph_output.pbtxt`, freezing_graph_pnn.py:129-139):
// fills `out` from its float `Const` nodes and infers width / kind from their names and shapes.  Throws on malformed input.
void read_frozen_graph(const std::vector<char>& data, const std::string& path, FlatFile* out);

}  // namespace pnn
