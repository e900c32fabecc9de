"""hm/direct/patch_hm.py on miniature stand-ins of the reference's TEncSearch.cpp fast pass (the real file lives in
/root/reference and is patched at build time only): the inserted prefetch call, the mode order 0..17, 19..34, 18 with the
candidate list updated in the original order (substitution), nothing reordered when the neural-network mode is outside the
loop (switch), and loud failures when an anchor is missing."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('patch_hm', os.path.join(ROOT, 'hm', 'direct', 'patch_hm.py'))
patch_hm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(patch_hm)

SEARCH = '''    bool contextFlag(false);
    initIntraPatternChType(tuRecurseWithPU,
                           contextFlag,
                           COMPONENT_Y,
                           true DEBUG_STRING_PASS_INTO(sTemp2));
    Bool doFastSearch = (numModesForFullRD != numModesAvailable);
    if (doFastSearch)
    {
      for (Int modeIdx(0); modeIdx < numModesAvailable; modeIdx++)
      {
        UInt uiMode(modeIdx);
        Double cost(evaluate(uiMode));
        CandNum += xUpdateCandList(uiMode,
                                   cost,
                                   numModesForFullRD,
                                   uiRdModeList,
                                   CandCostList);
      }
    }
'''


def test_prefetch_is_inserted_once_after_the_context_extraction():
    text, where = patch_hm.patch_search_cpp(SEARCH)
    assert text.count('pnn_hm_direct::prefetch(m_pnn') == 1
    assert text.index('Bool doFastSearch') < text.index('pnn_hm_direct::prefetch') < text.index('if (doFastSearch)')
    assert text[where[0]:where[0] + where[1]].startswith('      for (Int modeIdx(0)')


def test_substitution_loop_keeps_the_list_update_order():
    text, where = patch_hm.patch_search_cpp(SEARCH)
    out = patch_hm.reorder_mode_loop(text, where, 18)
    assert 'modeOrder < 18 ? modeOrder : (modeOrder < numModesAvailable - 1 ? modeOrder + 1 : 18)' in out
    assert out.count('xUpdateCandList(') == 1 and 'costOfMode[modeIdx] = cost;' in out
    assert out.index('costOfMode[modeIdx] = cost;') < out.index('xUpdateCandList(static_cast<UInt>(modeIdx)') < out.index('prefetch_first_quadrant(')
    assert patch_hm.add_post_loop(text, where).count('prefetch_first_quadrant(') == 1
    # the visiting order the patched loop produces, and the order in which the list is updated
    order = [m if m < 18 else (m + 1 if m < 34 else 18) for m in range(35)]
    assert sorted(order) == list(range(35)) and order[-1] == 18 and order[:18] == list(range(18))


def test_mode_index_detection_and_missing_anchors():
    assert patch_hm.nn_mode_of_loop('...if (uiDirMode != 18)\n...') == 18
    assert patch_hm.nn_mode_of_loop('...if (uiDirMode != 35)\n...') is None
    with pytest.raises(SystemExit):
        patch_hm.nn_mode_of_loop('no such test')
    with pytest.raises(SystemExit):
        patch_hm.patch_search_cpp(SEARCH.replace('Bool doFastSearch', 'Bool fast'))
    with pytest.raises(SystemExit):
        patch_hm.patch_search_cpp(SEARCH.replace('initIntraPatternChType(tuRecurseWithPU', 'somethingElse(tuRecurseWithPU'))
    with pytest.raises(SystemExit):
        patch_hm.patch_search_cpp(SEARCH.replace('CandCostList);', 'CandCosts);'))


@pytest.mark.skipif(not os.path.isdir('/root/reference/hevc'), reason='the reference tree is only present in the build container')
@pytest.mark.parametrize('variant, nn_mode', [('substitution', 18), ('switch', None)])
def test_patch_applies_to_the_reference_tree(variant, nn_mode, tmp_path):
    lib = '/root/reference/hevc/hm_16_15_%s/source/Lib' % variant
    text = open(os.path.join(lib, 'TLibEncoder', 'TEncSearch.cpp'), encoding='latin-1').read()
    patched, where = patch_hm.patch_search_cpp(text)
    assert patch_hm.nn_mode_of_loop(open(os.path.join(lib, 'TLibCommon', 'TComPrediction.cpp'), encoding='latin-1').read()) == nn_mode
    if nn_mode is not None:
        patched = patch_hm.reorder_mode_loop(patched, where, nn_mode)
        assert patched.count('costOfMode[modeIdx]') == 2
    else:
        patched = patch_hm.add_post_loop(patched, where)
    assert len(patched) > len(text) and patched.count('pnn_hm_direct::prefetch(') == 1
    # the request of the top-left quadrant is posted after the mode loop, before the most-probable-mode handling
    assert patched.count('pnn_hm_direct::prefetch_first_quadrant(') == 1
    assert patched.index('xUpdateCandList(') < patched.index('prefetch_first_quadrant(') < patched.index('getIntraDirPredictor(uiPartOffset')
