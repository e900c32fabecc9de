// Direct binding of the reference's HM hooks to libpnn_cuda.
//
// hm/direct/patch_hm.py rewrites the three hook regions of the reference's TComPrediction.{h,cpp} and TComPattern.cpp
// (copies made at build time under /tmp, never stored in this repository) so that they call the functions below instead
// of TensorFlow / embedded Python:
//   TComPrediction::initTempBuff   (TComPrediction.cpp(substitution):108-236)  -> read_mean_file + create
//   initIntraPatternChType         (TComPattern.cpp:342-380)                   -> set_context   (pnn_set_context)
//   predIntraAng, NN branch        (TComPrediction.cpp:556-635)                -> predict       (pnn_predict_hm)
//   estIntraPredLumaQT, fast pass  (TEncSearch.cpp:2332-2342), one inserted call  -> prefetch      (pnn_predict_hm_begin)
// Unlike the link seam of hm/shim/ (which only replaces Session::Run), the context gather / masking / mean subtraction
// and the add-mean / clip / round epilogue run inside the library here.
#ifndef PNN_HM_DIRECT_H
#define PNN_HM_DIRECT_H

#include <string>

#include "pnn_cuda.h"

namespace pnn_hm_direct {

// The reference's mean file: a pickled Python float (protocol 0 / 1 / 2, sets/results/training_set/means/luminance/
// mean_training.pkl) or a text float.
float read_mean_file(const std::string& path_to_mean_training);

// pnn_create on the paths file (`width,is_pair,0,path`, single / pair models by QP >= 32) + the settings an HM process
// wants: result memo on, lazy context staging on; per-width call statistics are written at exit (PNN_HM_STATS=<file>).
pnn_handle* create(const std::string& path_to_file_paths_to_graphs_output, float mean_training, int qp_selection);

// extract_context_portions' arguments (extraction_context.h:36-48) without the destination buffers
int set_context(pnn_handle* handle, int width, const int* piRoiOrigin, int iPicStride, const bool* bNeighborFlags,
                int iNumIntraNeighbor, int iUnitWidth, int iUnitHeight, int iAboveUnits, int iLeftUnits);

// Fast pass of the intra search (TEncSearch.cpp(substitution):2332-2393), right after initIntraPatternChType: the context
// of the PU is final, the neural-network mode is evaluated further down (substitution: mode 18 of the loop; switch: mode 35
// of the RD list) -> the request is posted now (pnn_predict_hm_begin) and computes while the host predicts and costs the
// other modes.  PNN_HM_PREFETCH=0 turns it off (A/B measurements with one executable).
int prefetch(pnn_handle* handle, int width);

// After the mode loop of the fast pass: the same for the top-left quadrant of the PU (initIntraPatternChType has just been run
// on it), whose context is final already; the answer waits in the memo until the codec reaches that PU.
// PNN_HM_PREFETCH_QUADRANT=0 turns it off.
int prefetch_first_quadrant(pnn_handle* handle, int width);

// NN branch of predIntraAng: prediction of the staged context into HM's Pel buffer
int predict(pnn_handle* handle, int width, int* piPred, int stride);

}  // namespace pnn_hm_direct

#endif
