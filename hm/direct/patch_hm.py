"""Writes patched copies of the three hook files of a reference HM variant.

    python hm/direct/patch_hm.py <reference source/Lib/TLibCommon> <output directory>

TComPrediction.h, TComPrediction.cpp and TComPattern.cpp are read from the reference tree, the TensorFlow / embedded
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


def main():
    src, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    for name, fn in (('TComPrediction.h', patch_prediction_h), ('TComPrediction.cpp', patch_prediction_cpp),
                     ('TComPattern.cpp', patch_pattern_cpp)):
        text = open(os.path.join(src, name), encoding='latin-1').read()
        open(os.path.join(out, name), 'w', encoding='latin-1').write(fn(text))
    print('patched TComPrediction.h, TComPrediction.cpp, TComPattern.cpp -> ' + out)


if __name__ == '__main__':
    main()
