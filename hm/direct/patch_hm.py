"""Writes patched copies of the three hook files of a reference HM variant.

    python hm/direct/patch_hm.py <reference source/Lib/TLibCommon> <output directory> [<reference source/Lib/TLibEncoder> <output directory>]

TComPrediction.h, TComPrediction.cpp and TComPattern.cpp (and, with the optional pair of directories, TEncSearch.cpp, which
gets one inserted call) are read from the reference tree, the TensorFlow / embedded
Python regions are replaced by calls into hm/direct/pnn_hm_direct.h (libpnn_cuda), and the result is written to the
output directory (under /tmp at build time: reference sources are never stored in this repository).  Every region is
located by anchors that exist exactly once; the script fails loudly if the reference text changes.
"""
import os
import sys


def cut(text, start, end, replacement, name, after=0):
    """Replaces text[start anchor .. end anchor] (both inclusive); each anchor must be found."""
    i = text.find(start, after)
    if i < 0:
        raise SystemExit('patch_hm: start anchor of `%s` not found' % name)
    j = text.find(end, i)
    if j < 0:
        raise SystemExit('patch_hm: end anchor of `%s` not found' % name)
    return text[:i] + replacement + text[j + len(end):]


def patch_prediction_h(text):
    text = cut(text, '#include "integration_prediction_neural_network.h"\n', '#include "interface_c_python.h"\n',
               '#include <map>\n#include <memory>\n#include <string>\n#include <vector>\n#include "pnn_hm_direct.h"\n', 'includes')
    start = '    std::vector<std::unique_ptr<tensorflow::Session>> m_vector_unique_ptrs_session;'
    end_line = 'm_tensors_portion_left;'
    i = text.find(start)
    j = text.find(end_line, i)
    if i < 0 or j < 0:
        raise SystemExit('patch_hm: TensorFlow members not found')
    j = text.find('\n', j) + 1
    return text[:i] + '    pnn_handle* m_pnn = NULL;   /**< libpnn_cuda handle: the five prediction neural networks. */\n' + text[j:]


def patch_prediction_cpp(text):
    text = cut(text, '    /*\n    The memory for the tensors storing the masked contexts', '    Py_Finalize();\n', '''    // libpnn_cuda: replaces create_tensors_*, the parse of the paths file, the single / pair choice by QP, load_graphs and
    // the embedded-Python pickle loader
    m_meanTraining = pnn_hm_direct::read_mean_file(path_to_mean_training);
    if (m_pnn == NULL)
    {
        m_pnn = pnn_hm_direct::create(path_to_file_paths_to_graphs_output, m_meanTraining, qp_selection);
        if (m_pnn == NULL)
        {
            assert(false);
        }
    }
''', 'initTempBuff')
    text = cut(text, '                /*\n                A fully-connected prediction neural network is', '                    pDstTemp += uiStride;\n                }\n',
               '''                // libpnn_cuda: net selected by the width, add-mean / clip / round epilogue fused, HM stride
                const int error_code_pnn(pnn_hm_direct::predict(m_pnn, iWidth, pDst, static_cast<int>(uiStride)));
                if (error_code_pnn < 0)
                {
                    assert(false);
                }
''', 'predIntraAng')
    if 'tensorflow' in text or 'Py_' in text:
        raise SystemExit('patch_hm: TensorFlow / Python references are left in TComPrediction.cpp')
    return text


def patch_pattern_cpp(text):
    text = cut(text, '        float* piPortionAbove(NULL);', '        assert(error_code >= 0);\n', '''        if (bitDepthForChannel != 8)
        {
            std::cerr << "The neural networks mode is not coded for bitdepths different from 8." << std::endl;
            assert(false);
        }
        // libpnn_cuda: replaces extract_context_portions into the TensorFlow input tensors
        const int error_code(pnn_hm_direct::set_context(m_pnn, static_cast<int>(uiTuWidth), piRoiOrigin, iPicStride, bNeighborFlags,
                                                        iNumIntraNeighbor, iUnitWidth, iUnitHeight, iAboveUnits, iLeftUnits));
        if (error_code < 0)
        {
            assert(false);
        }
''', 'initIntraPatternChType')
    # the optional overlap check read the float portions, which no longer exist on the host
    i = text.find('#ifdef CHECK_OVERLAP_INTRA_PATTERN_MASKED_CONTEXT_PORTIONS\n        error_code =')
    if i >= 0:
        j = text.find('#endif\n', i)
        text = text[:i] + text[j + len('#endif\n'):]
    return text


def patch_search_cpp(text):
    """TEncSearch::estIntraPredLumaQT, fast pass.

    1. One inserted call after initIntraPatternChType: the request of the neural-network mode is posted.
    2. After the mode loop: the request of the PU's top-left quadrant (the next PU the codec evaluates at this position) is posted.
    3. Substitution codec (the neural-network mode is number 18 of the loop): the loop visits the modes in the order
       0..17, 19..34, 18 and only stores their costs; the candidate list is then updated in the ORIGINAL order 0..34.  The
       iterations are independent (the prediction buffer is scratch, xModeBitsIntra reloads the entropy-coder state, the
       list update is the only carried state), so the list -- ties included -- and the bitstream are those of the
       unmodified loop, while the answer of the GPU is needed only after the host has evaluated the 34 other modes.
    """
    anchor = '    Bool doFastSearch = (numModesForFullRD != numModesAvailable);\n'
    if text.count(anchor) != 1:
        raise SystemExit('patch_hm: the fast-pass anchor of TEncSearch.cpp was found %d times' % text.count(anchor))
    i = text.find(anchor)
    before = text[max(0, i - 400):i]
    if 'initIntraPatternChType(tuRecurseWithPU' not in before or 'contextFlag' not in before:
        raise SystemExit('patch_hm: initIntraPatternChType does not precede the fast-pass anchor of TEncSearch.cpp')
    text = text[:i + len(anchor)] + '''    // libpnn_cuda: the context of this PU is final and the neural-network mode is evaluated further down (one of the
    // modes of the loop below, or of the RD list): post the request now, the other modes are evaluated meanwhile
    if (contextFlag)
    {
        const int error_code_prefetch(pnn_hm_direct::prefetch(m_pnn, static_cast<int>(tuRecurseWithPU.getRect(COMPONENT_Y).width)));
        if (error_code_prefetch < 0)
        {
            assert(false);
        }
    }
''' + text[i + len(anchor):]
    # does predIntraAng of this codec put the neural-network mode inside the 35 modes of the loop?
    loop_head = '      for (Int modeIdx(0); modeIdx < numModesAvailable; modeIdx++)\n      {\n        UInt uiMode(modeIdx);\n'
    update = '''        CandNum += xUpdateCandList(uiMode,
                                   cost,
                                   numModesForFullRD,
                                   uiRdModeList,
                                   CandCostList);
      }
'''
    j = text.find(loop_head, i)
    k = text.find(update, j)
    if j < 0 or k < 0 or text.count(loop_head) != 1:
        raise SystemExit('patch_hm: the mode loop of the fast pass was not found in TEncSearch.cpp')
    return text, (j, len(loop_head), k, len(update))


POST_LOOP = '''      // libpnn_cuda: the top-left quadrant of this PU is the first PU the codec looks at next at this position (first CU of the
      // next depth, or the first NxN PU); all of its context lies outside this PU, so it is final already: post its request
      // now, it computes during the RD pass of this PU and waits in the memo of in-loop results.  The memo is keyed by the
      // whole context, availability included: a context that turns out different later is simply computed again.
      if (contextFlag && puRect.width >= 8)
      {
        TComTURecurse tuFirstQuadrant(tuRecurseWithPU, false, TComTU::QUAD_SPLIT);
        bool contextFlagFirstQuadrant(false);
        initIntraPatternChType(tuFirstQuadrant,
                               contextFlagFirstQuadrant,
                               COMPONENT_Y,
                               true DEBUG_STRING_PASS_INTO(sTemp2));
        if (contextFlagFirstQuadrant)
        {
          const int error_code_prefetch(pnn_hm_direct::prefetch_first_quadrant(m_pnn, static_cast<int>(puRect.width / 2)));
          if (error_code_prefetch < 0)
          {
            assert(false);
          }
        }
      }
'''


def add_post_loop(text, where):
    j, n_head, k, n_update = where
    return text[:k + n_update] + POST_LOOP + text[k + n_update:]


def reorder_mode_loop(text, where, nn_mode):
    j, n_head, k, n_update = where
    head = '''      // libpnn_cuda: modes in the order 0..%d, %d..34, %d (the posted neural-network request is collected last); the costs
      // are stored and the candidate list is updated afterwards in the original order, see hm/direct/patch_hm.py
      Double costOfMode[35];
      assert(numModesAvailable == 35);
      for (Int modeOrder(0); modeOrder < numModesAvailable; modeOrder++)
      {
        const Int modeIdx(modeOrder < %d ? modeOrder : (modeOrder < numModesAvailable - 1 ? modeOrder + 1 : %d));
        UInt uiMode(modeIdx);
''' % (nn_mode - 1, nn_mode + 1, nn_mode, nn_mode, nn_mode)
    tail = '''        costOfMode[modeIdx] = cost;
      }
      for (Int modeIdx(0); modeIdx < numModesAvailable; modeIdx++)
      {
        CandNum += xUpdateCandList(static_cast<UInt>(modeIdx),
                                   costOfMode[modeIdx],
                                   numModesForFullRD,
                                   uiRdModeList,
                                   CandCostList);
      }
'''
    return text[:j] + head + text[j + n_head:k] + tail + POST_LOOP + text[k + n_update:]


def nn_mode_of_loop(prediction_cpp_text):
    """18 for the substitution codec (`if (uiDirMode != 18)` guards the regular modes), None when the neural-network mode
    lies outside the 35 modes of the fast pass (switch codec: 35)."""
    if 'if (uiDirMode != 18)' in prediction_cpp_text:
        return 18
    if 'if (uiDirMode != 35)' in prediction_cpp_text:
        return None
    raise SystemExit('patch_hm: the index of the neural-network mode was not found in TComPrediction.cpp')


def main():
    src, out = sys.argv[1], sys.argv[2]
    if len(sys.argv) > 4:      # <reference source/Lib/TLibEncoder> <output directory>: the fast-pass prefetch
        os.makedirs(sys.argv[4], exist_ok=True)
        text = open(os.path.join(sys.argv[3], 'TEncSearch.cpp'), encoding='latin-1').read()
        text, where = patch_search_cpp(text)
        nn_mode = nn_mode_of_loop(open(os.path.join(src, 'TComPrediction.cpp'), encoding='latin-1').read())
        text = reorder_mode_loop(text, where, nn_mode) if nn_mode is not None else add_post_loop(text, where)
        open(os.path.join(sys.argv[4], 'TEncSearch.cpp'), 'w', encoding='latin-1').write(text)
        print('patched TEncSearch.cpp (prefetch%s) -> %s' % ('' if nn_mode is None else ', neural-network mode %d evaluated last' % nn_mode, sys.argv[4]))
    os.makedirs(out, exist_ok=True)
    for name, fn in (('TComPrediction.h', patch_prediction_h), ('TComPrediction.cpp', patch_prediction_cpp),
                     ('TComPattern.cpp', patch_pattern_cpp)):
        text = open(os.path.join(src, name), encoding='latin-1').read()
        open(os.path.join(out, name), 'w', encoding='latin-1').write(fn(text))
    print('patched TComPrediction.h, TComPrediction.cpp, TComPattern.cpp -> ' + out)


if __name__ == '__main__':
    main()
