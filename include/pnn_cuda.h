/*
 * libpnn_cuda -- C ABI of the B200-native prediction-neural-network (PNN) engine.
 *
 * Drop-in boundary for the PNN forward pass of
 * thierrydumas/context_adaptive_neural_network_based_prediction.  Every entry point
 * cites the reference interface it replaces (paths relative to the reference root).
 * Plain pointers and sizes only; no torch / TensorFlow types.  All functions return
 * 0 on success and -1 on error (the reference's hm_common convention,
 * hevc/hm_common/c++/source_common/extraction_context.cpp:17-47); the message is
 * available from pnn_last_error().  A handle is NOT thread-safe (the reference keeps
 * one TComPrediction per HM process).
 *
 * There is no CPU fallback: pnn_create fails when no sm_100 device is present.
 */
#ifndef PNN_CUDA_H
#define PNN_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pnn_handle pnn_handle;

/* Arithmetic used by the GEMM-shaped layers. */
enum {
    PNN_PRECISION_FP32 = 0,    /* fp32 FFMA kernels: the accuracy yard-stick                       */
    PNN_PRECISION_BF16X3 = 1   /* tcgen05 tensor cores, operands split hi/lo bf16, 3 MMAs, fp32 acc */
};

/* Output rounding of the fused epilogue. */
enum {
    PNN_ROUND_HALF_EVEN = 0,   /* numpy.round, reference tools/tools.py:49                         */
    PNN_ROUND_HALF_AWAY = 1    /* std::round, reference TComPrediction.cpp(substitution):632       */
};

/*
 * Replaces TComPrediction::initTempBuff's network set-up
 * (hevc/hm_16_15_substitution/source/Lib/TLibCommon/TComPrediction.cpp:108-236) and
 * load_graphs (hevc/hm_common/c++/source_common/integration_prediction_neural_network.cpp:29-69).
 *
 * `paths_file` is the reference's "paths to graphs" file (lines `width,is_pair,0,path`, e.g.
 * hevc/hm_common/paths_to_graphs_output/pair.txt) whose paths now name PNNW flat binaries instead
 * or, unchanged, the frozen graphs themselves (see pnn_load_net); it may be NULL, in which case nets are added with pnn_load_net.
 * `qp_selection` selects the "pair" models when it is >= 32 and the file lists pair entries
 * (TComPrediction.cpp:156); it must be > 0 (TComPrediction.cpp:129-133).
 * `mean_training` is the training-set mean (TComPrediction.cpp:219), `device` the CUDA ordinal.
 */
int pnn_create(const char* paths_file, float mean_training, int qp_selection, int device, pnn_handle** out);

/*
 * Same, but nothing touches the GPU yet: the paths file and the headers of the weight files are validated at once, the
 * device check, the CUDA context and every buffer wait for the first call that needs them (which then fails loudly if there
 * is no sm_100 device).  For short-lived processes: a decoder whose bitstream never uses the neural-network mode finishes
 * without paying the ~1-2 s of CUDA start-up.  What the HM bindings use.
 */
int pnn_create_deferred(const char* paths_file, float mean_training, int qp_selection, int device, pnn_handle** out);

/*
 * After pnn_create_deferred: starts the device initialisation and the upload of the registered nets on a thread of the
 * library, so that they overlap the caller's own start-up (an encoder works through its first row of coding tree units,
 * whose blocks have no causal context, before it asks for the first prediction).  Calls that need the device wait for it.
 */
int pnn_warm_up(pnn_handle* h);

/* Replaces std::unique_ptr<tensorflow::Session> teardown. */
void pnn_destroy(pnn_handle* h);

/*
 * For a process that is about to exit: stops the persistent kernel and waits for the device, but frees nothing (releasing
 * hundreds of device buffers one by one costs seconds; process exit returns them at once).  The handle must not be used
 * afterwards.
 */
int pnn_release_at_exit(pnn_handle* h);

/* Last error message of the handle (or of the failed pnn_create when h is NULL). */
const char* pnn_last_error(pnn_handle* h);

/*
 * Loads the weights of one net -- replaces one load_graph call
 * (integration_prediction_neural_network.cpp:29-54) / PredictionNeuralNetwork.initialization
 * (pnn/PredictionNeuralNetwork.py:185-200).  `path` names either a PNNW flat binary (width and type in
 * its header) or the frozen graph the reference's HM loads (the binary GraphDef freezing_graph_pnn.py:129-139
 * writes as `graph_output.pbtxt`): its float `Const` nodes carry the variable names, width and type are
 * inferred from them.  No TensorFlow / protobuf library is involved.
 */
int pnn_load_net(pnn_handle* h, const char* path);

/*
 * Same as pnn_load_net, but only VALIDATES the file now (a PNNW header listing every tensor the net needs with the right
 * shapes and sizes, or a whole frozen graph) and uploads the weights the first time a call needs this net.  A decoder that
 * never meets the neural-network mode at some block width never pays for that net (transform blocks are <= 32x32,
 * hevc/configuration/intra_main_rext.cfg:13, so a decoder never uses the 64x64 net).  pnn_create with a paths file
 * registers its five nets this way.
 */
int pnn_register_net(pnn_handle* h, const char* path);

/*
 * Host-only (needs no GPU): parses a weights file exactly as pnn_load_net does and reports what it holds.
 * `checksum` = sum over the tensors in name order of sum_i (i % 7 + 1) * value_i, in double.  Any output may be NULL
 * (without `checksum` only the header of a PNNW file is read).
 * Returns 0, or -1 with the message in pnn_last_error(NULL).
 */
int pnn_inspect_net_file(const char* path, int* width_target, int* is_fully_connected, int64_t* n_parameters,
                         double* checksum);

/* PNN_PRECISION_*; default PNN_PRECISION_BF16X3. */
int pnn_set_precision(pnn_handle* h, int precision);

/*
 * Step 1 of the in-loop call: mirrors extract_context_portions
 * (hevc/hm_common/c++/source_common/extraction_context.h:36-48, called from
 * TComPattern.cpp(substitution):366-380).  `roi_origin` points at the top-left pixel of the
 * current TB in HM's int reconstruction plane (row stride `pic_stride`); `neighbor_flags` holds
 * left_units + above_units + 1 availability flags ordered bottom-left -> top-left, above-left,
 * above -> above-right (TComPattern.cpp:260-280).  The context pixels are staged; masking and
 * mean subtraction run on the device inside pnn_predict_hm.  A second call overwrites the first.
 */
int pnn_set_context(pnn_handle* h, int width, const int32_t* roi_origin, int pic_stride,
                    const uint8_t* neighbor_flags, int num_intra_neighbor,
                    int unit_width, int unit_height, int above_units, int left_units);

/*
 * 1: pnn_set_context only records its arguments (after the same checks) and pnn_predict_hm copies the context pixels,
 * so a context that is never predicted from costs nothing -- HM extracts one for every transform block and every
 * candidate mode (TEncSearch.cpp(substitution):1229-1235), the neural-network mode needs few of them.  The caller then
 * guarantees that the reconstruction plane and `roi_origin` stay unchanged between the two calls, which holds in HM
 * (prediction precedes reconstruction).  0 (default): pixels are copied by pnn_set_context.
 */
int pnn_set_context_lazy(pnn_handle* h, int enabled);

/*
 * Step 2 of the in-loop call: the neural-network branch of TComPrediction::predIntraAng
 * (TComPrediction.cpp(substitution):556-635): runs the net selected by `width` on the staged
 * context and writes (int)std::round(clip(p + mean, 0, 255)) into dst[row*dst_stride + col].
 * Synchronous: dst is complete on return.
 */
int pnn_predict_hm(pnn_handle* h, int width, int32_t* dst, int dst_stride);

/*
 * Optional step 1b: posts the context of the last pnn_set_context to the GPU and returns at once, so that the codec's own
 * work hides the latency of the call.  In the fast pass of the intra search (TEncSearch.cpp(substitution):2332-2393) the
 * context is final once initIntraPatternChType has run, the neural-network mode is number 18 of the 35 the loop
 * evaluates: the request travels and computes while the host predicts and costs modes 0..17.  A pnn_predict_hm of the
 * same width that follows without another pnn_set_context collects the answer; any other call first finishes the request
 * and keeps its answer in the memo of in-loop results (pnn_set_hm_cache), where the RD pass of the switch codec -- which
 * evaluates mode 35 for the very transform block the fast pass just looked at (TEncSearch.cpp(switch):2476-2491) -- finds
 * it.  Nothing is speculative: the context is the final one, the bits are those of pnn_predict_hm.  One request in flight.
 */
int pnn_predict_hm_begin(pnn_handle* h, int width);

/*
 * In-loop call with an already extracted context: what Session::Run does in the reference's HM
 * (TComPrediction.cpp(substitution):572-579 / 601-608) after extract_context_portions filled the
 * batch-1 input tensors (TComPattern.cpp:366-380).  FC nets: `above_or_flat` is [5*W*W], `left` NULL;
 * conv nets: `above_or_flat` [W*3W], `left` [2W*W] (mean-centred, masked floats).  `out` receives the raw
 * float32 prediction [W*W] like the output tensor of Session::Run.  Same batch-1 kernels and reduction
 * order as pnn_predict_hm.  Synchronous.
 */
int pnn_predict_hm_context(pnn_handle* h, int width, const float* above_or_flat, const float* left, float* out);

/*
 * Offline path with already pre-processed contexts: pnn.batching.predict_by_batch_via_pnn
 * (pnn/batching.py:7-88).  FC nets: `above_or_flat` is [n, 5*W*W], `left` is NULL.
 * Conv nets: `above_or_flat` is [n, W, 3W, 1] and `left` is [n, 2W, W, 1].  `out` receives the raw
 * float32 predictions [n, W, W, 1] (mean not added, like sess.run).  HOST pointers.
 */
int pnn_predict_batch(pnn_handle* h, int width, int is_fully_connected,
                      const float* above_or_flat, const float* left, int64_t n, float* out);

/*
 * Offline path with the gather fused on the device: replaces
 * extract_context_portions_targets_from_channel_numpy + preprocess_context_portions_targets_numpy
 * (sets/common.py:13-263, 351-475), predict_by_batch_via_pnn and cast_float_to_uint8
 * (tools/tools.py:17-49) for blocks of `n_images` uint8 images [n_images, height, width_image].
 * Block i is the W x W target whose top-left pixel is (rows[i], cols[i]) of image image_index[i]
 * (image_index may be NULL when n_images == 1).  mask_w / mask_h in {0,4,...,W} as in
 * sets/common.py:444-447; context pixels outside the image are masked too.
 * Outputs (each may be NULL): out_f32 raw predictions [n, W*W]; out_u8 =
 * round_half_even(clip(p + mean, 0, 255)) [n, W*W]; out_psnr[n] = PSNR (float64, tools/tools.py:364-401)
 * between the target block and out_u8.  HOST pointers.
 */
int pnn_predict_image_blocks(pnn_handle* h, int width, int is_fully_connected,
                             const uint8_t* images, int n_images, int height, int width_image,
                             const int32_t* image_index, const int32_t* rows, const int32_t* cols, int64_t n,
                             int mask_w, int mask_h, float* out_f32, uint8_t* out_u8, double* out_psnr);

/*
 * Asynchronous form of pnn_predict_image_blocks for callers that evaluate several block sizes (or several image sets)
 * back to back, as comparing_pnn_ipfcns_hevc_best_mode.py:220-322 does per width: the call returns once the work is
 * enqueued; uploads, kernels and read-backs run on three streams, so the upload of call k+1 and the read-back of call k
 * overlap the kernels of the other call.  The host buffers (inputs AND outputs) must stay valid, and should be pinned,
 * until pnn_synchronize returns; outputs are complete only then.  Same arguments and checks as the synchronous call.
 */
int pnn_predict_image_blocks_async(pnn_handle* h, int width_target, int is_fully_connected, const uint8_t* images_uint8,
                                   int n_images, int height, int width_image, const int32_t* image_index,
                                   const int32_t* rows, const int32_t* cols, int64_t n, int mask_w, int mask_h,
                                   float* out_float32, uint8_t* out_uint8, double* out_psnr);

/* Waits for everything enqueued on the handle (pnn_predict_image_blocks_async). */
int pnn_synchronize(pnn_handle* h);

/*
 * Same two calls with DEVICE pointers, asynchronous on `cuda_stream` (a cudaStream_t, may be 0).
 * These are what the throughput numbers with inputs resident in HBM are measured on.
 * The block list lives on the device, so it cannot be validated here.  Precondition: every block and its context anchor lie
 * inside its image (what the host-pointer call checks).  A violation is contained, not undefined: context pixels outside the
 * image (or of an image index outside [0, n_images)) read as masked, and the PSNR of a block that leaves the image is NaN;
 * pnn_hevc_best_mode_device answers such a block with best_index 255, PSNR NaN and a zero prediction.
 */
int pnn_predict_batch_device(pnn_handle* h, int width, int is_fully_connected,
                             const float* d_above_or_flat, const float* d_left, int64_t n, float* d_out,
                             void* cuda_stream);
int pnn_predict_image_blocks_device(pnn_handle* h, int width, int is_fully_connected,
                                    const uint8_t* d_images, int n_images, int height, int width_image,
                                    const int32_t* d_image_index, const int32_t* d_rows, const int32_t* d_cols,
                                    int64_t n, int mask_w, int mask_h,
                                    float* d_out_f32, uint8_t* d_out_u8, double* d_out_psnr, void* cuda_stream);

/*
 * In-loop latency path.  1 (default): the FC nets of widths 4 and 8 are served by ONE persistent kernel whose CTAs keep
 * the weights of both nets in shared memory for the life of the kernel; a call is a doorbell write into mapped pinned
 * memory and a poll of the answer, no launch.  The kernel leaves the SMs whenever the handle is asked for anything else
 * (a convolutional net, the offline path) and comes back behind that work.  Convolutional nets: CUDA graph with split-K.
 * 0: every net is a CUDA graph of plain launches (FC: one launch per layer).  Both settings give identical bits for the
 * FC nets (same device functions, same reduction order); if the device cannot hold the persistent kernel (148 co-resident
 * CTAs) the library falls back to 0 for the FC nets by itself.
 */
int pnn_set_hm_fused(pnn_handle* h, int enabled);

/*
 * Memo of in-loop results (off by default; the HM bindings switch it on): the codec evaluates the same prediction unit
 * with the same context several times (fast pass, rate-distortion pass, final reconstruction:
 * TEncSearch.cpp(substitution):2331-2393, 1229-1250), and a PNN is a pure function of its context, so a context seen
 * before is answered from host memory -- bit-identical by construction (the whole context is compared, not a hash).
 */
int pnn_set_hm_cache(pnn_handle* h, int enabled);
int pnn_hm_cache_stats(pnn_handle* h, int64_t* hits, int64_t* misses);

/*
 * Bytes of activation workspace one net may use (default 20 GB, or PNN_WORKSPACE_GB at pnn_create): batched calls are
 * cut into chunks that fit.  Results do not depend on it (a block's prediction is independent of its batch).
 * The budget is an upper bound, not a reservation: a workspace grows to what the largest call so far needed, and when the
 * device is short of memory (another engine or framework on the same GPU) the library first gives back the workspaces of
 * the handle's other nets, then halves the chunk until it fits.
 */
int pnn_set_workspace_budget(pnn_handle* h, int64_t bytes_per_net);

/* Kernels launched by this handle since creation (bench.py's gpu_launches). */
int64_t pnn_launch_count(pnn_handle* h);

/* Device time (ms, CUDA events on the handle's stream) of the last pnn_predict_hm call. */
float pnn_last_hm_device_ms(pnn_handle* h);

/*
 * Baseline of the offline evaluation: the best of the 35 HEVC intra prediction modes for every block, replacing
 * extract_intra_patterns + predict_series_via_hevc_best_mode (hevc/intraprediction/intraprediction.py:8-292, which loops
 * over hevc_intraprediction, hevc/intraprediction/c++/source/extracted_hevc_intraprediction.cpp:3-421, through Cython).
 * Blocks are addressed like in pnn_predict_image_blocks; the intra pattern starts at (row - 1, col - 1)
 * (comparing_pnn_ipfcns_hevc_best_mode.py:234-235), is 2W + 1 - mask_h high and 2W + 1 - mask_w wide, and its missing
 * part (masked or outside the image) is padded with the last pixel as the reference does.  Integer arithmetic, bit-exact.
 * Outputs (each may be NULL): best_index[n] (first mode with the highest PSNR), psnr[n] float64, pred_u8 [n, W*W].
 * The first variant takes HOST pointers, the second DEVICE pointers (asynchronous on `cuda_stream`).
 */
int pnn_hevc_best_mode(pnn_handle* h, int width, const uint8_t* images, int n_images, int height, int width_image,
                       const int32_t* image_index, const int32_t* rows, const int32_t* cols, int64_t n, int mask_w, int mask_h,
                       uint8_t* best_index, double* psnr, uint8_t* pred_u8);
int pnn_hevc_best_mode_device(pnn_handle* h, int width, const uint8_t* d_images, int n_images, int height, int width_image,
                              const int32_t* d_image_index, const int32_t* d_rows, const int32_t* d_cols, int64_t n,
                              int mask_w, int mask_h, uint8_t* d_best_index, double* d_psnr, uint8_t* d_pred_u8,
                              void* cuda_stream);

/*
 * Win flags of the offline evaluation (reference comparing_pnn_ipfcns_hevc_best_mode.py:87):
 * d_win[i] = (d_psnr[i] - d_psnr_baseline[i] > 0).  DEVICE pointers, asynchronous on `cuda_stream`.
 */
int pnn_win_flags_device(pnn_handle* h, const double* d_psnr, const double* d_psnr_baseline, int64_t n,
                         uint8_t* d_win, void* cuda_stream);

/*
 * Per-kernel device timing (CUDA events around every launch on the launching stream).  Off by default.
 * pnn_profile_report synchronises, aggregates by (kernel, shape) and returns a text table
 * "kernel M N K launches ms flops"; the totals of the tcgen05 / fp32 GEMM kernel are returned in
 * *gemm_ms / *gemm_flops / *gemm_launches and those of all other kernels in *other_ms (each may be NULL).
 * The report resets the counters.
 */
int pnn_set_profiling(pnn_handle* h, int enabled);
const char* pnn_profile_report(pnn_handle* h, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches,
                               double* other_ms);

/*
 * Inspection hook (tracing aid, no reference counterpart): copies activation buffer `buffer_index` of the
 * given net, as left by the LAST call, to the host as float32 [n_samples, elements per sample].
 * Buffer 0.. follow the layer order; returns the elements per sample in *elems_per_sample
 * (out may be NULL to query it).
 */
int pnn_debug_get_activation(pnn_handle* h, int width, int is_fully_connected, int buffer_index, int64_t n_samples,
                             float* out, int64_t* elems_per_sample);

/*
 * Tuning aid (no reference counterpart): times `iters` launches of the tcgen05 GEMM kernel on a synthetic
 * fully-connected problem out[M,N] = lrelu(in[M,K] * W[K,N] + b) with zero-filled operands and returns the
 * mean device time per launch in milliseconds (< 0 on error).  `flags`: bit 0 skip the A (activation) loads,
 * bit 1 skip the B (weight) loads, bit 2 skip the epilogue stores -- results are then meaningless, only the
 * time is of interest.
 */
float pnn_debug_time_gemm(pnn_handle* h, int64_t M, int N, int K, int iters, int flags);

/* Library version string. */
const char* pnn_version(void);

#ifdef __cplusplus
}
#endif

#endif /* PNN_CUDA_H */
