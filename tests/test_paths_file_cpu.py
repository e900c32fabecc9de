"""a12: the paths file of TComPrediction::initTempBuff (`width,is_pair,0,path`, hevc/hm_common/paths_to_graphs_output/*.txt).
`pnn_create[_deferred]` must choose, for every width, the path the REFERENCE's own parser and selection choose
(tools.cpp:52-110 `parse_file_strings_three_keys`, compiled unmodified into oracle/_ref/libtools_ref.so, +
TComPrediction.cpp(substitution):145-171) -- on plain files and on the awkward ones: blanks around fields, CRLF, runs of
delimiters, a non-zero third key, duplicates, blank lines, no final newline.  No GPU: creation is deferred."""
import ctypes
import os
import shutil

import pytest

import helpers
from context_adaptive_neural_network_based_prediction_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WIDTHS = (4, 8, 16, 32, 64)

CASES = {
    'plain_single': '4,0,0,{s4}\n8,0,0,{s8}\n16,0,0,{s16}\n32,0,0,{s32}\n64,0,0,{s64}\n',
    'single_and_pair': ''.join('%d,0,0,{s%d}\n%d,1,0,{p%d}\n' % (w, w, w, w) for w in WIDTHS),
    'blanks_and_crlf': ' 4 , 0 , 0 ,   {s4}  \r\n8,0,0,\t{s8}\t\r\n\r\n16,0,0,{s16}\r\n   \r\n32,0,0,{s32}\r\n64,0,0,{s64}',
    'runs_of_delimiters': '4,,0,,,0,{s4}\n8,0,0,,{s8}\n16,0,0,{s16},,\n32,0,0,{s32}\n64,0,0,{s64}\n',
    'third_key': '4,0,1,{x4}\n4,0,0,{s4}\n8,0,0,{s8}\n8,0,2,{x8}\n16,0,0,{s16}\n32,0,0,{s32}\n64,0,0,{s64}\n64,1,3,{y64}\n',
    'duplicates_last_wins': '4,0,0,{x4}\n4,0,0,{s4}\n8,0,0,{s8}\n16,0,0,{x16}\n16,0,0,{s16}\n32,0,0,{s32}\n64,0,0,{s64}\n',
    'pair_incomplete': ''.join('%d,0,0,{s%d}\n' % (w, w) for w in WIDTHS) + '4,1,0,{p4}\n8,1,0,{p8}\n',
    'width_missing': '4,0,0,{s4}\n8,0,0,{s8}\n32,0,0,{s32}\n64,0,0,{s64}\n',
}


@pytest.fixture(scope='module')
def nets(tmp_path_factory):
    d = str(tmp_path_factory.mktemp('paths_file_nets'))
    return {w: helpers.make_net_file(d, w, w <= 8, seed=w)[0] for w in WIDTHS}


@pytest.fixture(scope='module')
def reference_parser():
    path = os.path.join(ROOT, 'oracle', '_ref', 'libtools_ref.so')
    if not os.path.exists(path):
        pytest.skip('oracle/_ref/libtools_ref.so not built (needs /root/reference)')
    lib = ctypes.CDLL(path)
    lib.ref_choose_path.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_uint, ctypes.c_char_p, ctypes.c_int]
    return lib


def _create(lib, paths_file, qp):
    h = ctypes.c_void_p()
    code = lib.pnn_create_deferred(paths_file.encode(), ctypes.c_float(helpers.MEAN), qp, 0, ctypes.byref(h))
    if code == 0:
        lib.pnn_destroy(h)
    return code


@pytest.mark.parametrize('name', sorted(CASES))
@pytest.mark.parametrize('qp', [22, 32])
def test_library_chooses_what_the_reference_parser_chooses(name, qp, nets, reference_parser, tmp_path):
    lib = _lib.load()
    names = {}
    for kind in 'spxy':
        for w in WIDTHS:
            names['%s%d' % (kind, w)] = str(tmp_path / ('%s_%d.pnnw' % (kind, w)))
    paths_file = str(tmp_path / 'paths.txt')
    with open(paths_file, 'w', newline='') as f:
        f.write(CASES[name].format(**names))
    chosen = {}
    buf = ctypes.create_string_buffer(4096)
    for w in WIDTHS:
        code = reference_parser.ref_choose_path(paths_file.encode(), qp, w, buf, 4096)
        assert code >= 0
        chosen[w] = buf.value.decode() if code == 1 else None
    if any(v is None for v in chosen.values()):
        # the reference's map::at would throw here; the library reports it
        # (third_key: its `64,1,3,...` line makes the pair map non-empty, so QP 32 selects a map without any (width, 0) entry)
        assert (name, qp) in (('width_missing', 22), ('width_missing', 32), ('pair_incomplete', 32), ('third_key', 32))
        for w, path in chosen.items():
            if path is not None:
                shutil.copyfile(nets[w], path)
        assert _create(lib, paths_file, qp) == -1
        return
    # ONLY the files the reference chose exist: any other choice of the library would fail to open its file
    for w, path in chosen.items():
        assert path in names.values()
        shutil.copyfile(nets[w], path)
    assert _create(lib, paths_file, qp) == 0, lib.pnn_last_error(None)
    os.remove(chosen[16])
    assert _create(lib, paths_file, qp) == -1                      # (the check does discriminate)


def test_malformed_lines_are_reported(nets, tmp_path):
    """Where the reference's std::stoul / vector::at throw (and the codec aborts), the library returns -1."""
    lib = _lib.load()
    good = ''.join('%d,0,0,%s\n' % (w, nets[w]) for w in WIDTHS)
    for bad in ('4,0,0\n', 'four,0,0,%s\n' % nets[4], ',4,0,0,%s\n' % nets[4], '4,0,%s\n' % nets[4]):
        paths_file = str(tmp_path / 'bad.txt')
        open(paths_file, 'w').write(good + bad)
        assert _create(lib, paths_file, 22) == -1
    assert _create(lib, str(tmp_path / 'missing.txt'), 22) == -1
