#!/bin/bash
# Builds the reference's HM-16.15 "substitution" and "switch" codecs UNMODIFIED from /root/reference against
# libpnn_cuda through the link seam of hm/shim/ (two same-named headers first on the include path, see
# INTEGRATION.md).  Objects go to /tmp, only the four executables land in hm/_build/ (git-ignored, like
# oracle/_ref: they travel to the GPU box, the reference sources are never copied into this repository).
#   usage: hm/build_hm.sh [substitution|switch|regular ...]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${REF:-/root/reference}"
PKG="$ROOT/context_adaptive_neural_network_based_prediction_b200"
OUT="$ROOT/hm/_build"
JOBS="${JOBS:-8}"
mkdir -p "$OUT"
[ -f "$PKG/csrc/libpnn_cuda.so" ] || { echo "build libpnn_cuda.so first (__graft_entry__.build())"; exit 1; }
VARIANTS="${@:-substitution}"
for V in $VARIANTS; do
  SRC="$REF/hevc/hm_16_15_$V/source"
  COMMON="$REF/hevc/hm_common/c++/source_common"
  OBJ="/tmp/pnn_hm_build/$V"
  mkdir -p "$OBJ"
  # -include cmath: visualization_debugging.cpp uses std::round without <cmath> (gcc 13)
  FLAGS="-O3 -std=c++11 -DMSYS_LINUX -w -include cmath -I$ROOT/hm/shim -I$ROOT/include -I$SRC/Lib -I$SRC/Lib/TLibCommon -I$COMMON"
  LIBSRCS=$(ls $SRC/Lib/TLibCommon/*.cpp $SRC/Lib/TLibVideoIO/*.cpp $SRC/Lib/TLibEncoder/*.cpp $SRC/Lib/TLibDecoder/*.cpp $SRC/Lib/TAppCommon/*.cpp)
  if [ "$V" != "regular" ]; then
    LIBSRCS="$LIBSRCS $COMMON/extraction_context.cpp $COMMON/tools.cpp $COMMON/visualization_debugging.cpp $ROOT/hm/shim/pnn_hm_shim.cpp"
  else
    LIBSRCS="$LIBSRCS $COMMON/tools.cpp $COMMON/visualization_debugging.cpp"
  fi
  ENCSRCS=$(ls $SRC/App/TAppEncoder/*.cpp)
  DECSRCS=$(ls $SRC/App/TAppDecoder/*.cpp)
  compile() { # $1 = source, $2 = tag
    local o="$OBJ/$2_$(basename "${1%.*}").o"
    if [ ! -f "$o" ] || [ "$1" -nt "$o" ]; then g++ $FLAGS -c "$1" -o "$o" || exit 1; fi
  }
  export -f compile; export FLAGS OBJ
  printf '%s\n' $LIBSRCS | xargs -P "$JOBS" -I{} bash -c 'compile {} lib'
  printf '%s\n' $ENCSRCS | xargs -P "$JOBS" -I{} bash -c 'compile {} enc'
  printf '%s\n' $DECSRCS | xargs -P "$JOBS" -I{} bash -c 'compile {} dec'
  gcc -O3 -w -c "$SRC/Lib/libmd5/libmd5.c" -o "$OBJ/lib_libmd5.o"
  LINK="-lpthread -ldl"
  if [ "$V" != "regular" ]; then
    LINK="-L$PKG/csrc -lpnn_cuda -Wl,-rpath,\$ORIGIN/../../context_adaptive_neural_network_based_prediction_b200/csrc $LINK"
  fi
  g++ -o "$OUT/TAppEncoderStatic_$V" $OBJ/enc_*.o $OBJ/lib_*.o $LINK
  g++ -o "$OUT/TAppDecoderStatic_$V" $OBJ/dec_*.o $OBJ/lib_*.o $LINK
  echo "built $OUT/TAppEncoderStatic_$V and $OUT/TAppDecoderStatic_$V"
  # Baseline leg: the SAME objects linked against oracle/_ref/libpnn_ref.so, the libtorch-CPU
  # stand-in for the reference's TensorFlow-CPU build (same C ABI, same link seam) -> *_cpu executables
  if [ "$V" != "regular" ] && [ -f "$ROOT/oracle/_ref/libpnn_ref.so" ]; then
    LINKCPU="-L$ROOT/oracle/_ref -lpnn_ref -Wl,-rpath,\$ORIGIN/../../oracle/_ref -lpthread -ldl"
    g++ -o "$OUT/TAppEncoderStatic_${V}_cpu" $OBJ/enc_*.o $OBJ/lib_*.o $LINKCPU
    g++ -o "$OUT/TAppDecoderStatic_${V}_cpu" $OBJ/dec_*.o $OBJ/lib_*.o $LINKCPU
    echo "built $OUT/TAppEncoderStatic_${V}_cpu and $OUT/TAppDecoderStatic_${V}_cpu (CPU baseline backend)"
  fi
done
# Direct binding (hm/direct/): the three hook files are patched copies made under /tmp by
# hm/direct/patch_hm.py, everything else compiles from the reference tree as it lies -> *_direct executables
for V in $VARIANTS; do
  [ "$V" = "regular" ] && continue
  [ "${DIRECT:-1}" = "1" ] || continue
  SRC="$REF/hevc/hm_16_15_$V/source"
  COMMON="$REF/hevc/hm_common/c++/source_common"
  OBJ="/tmp/pnn_hm_build/${V}_direct"
  # a mirror of the source tree made of symbolic links (quote-includes resolve next to the including file, so the
  # patched header must sit among its neighbours), with the three hook files replaced by their patched copies
  MIRROR="$OBJ/mirror"
  rm -rf "$MIRROR"
  mkdir -p "$OBJ"
  cp -rs "$SRC" "$MIRROR"
  rm -f "$MIRROR/Lib/TLibCommon/TComPrediction.h" "$MIRROR/Lib/TLibCommon/TComPrediction.cpp" "$MIRROR/Lib/TLibCommon/TComPattern.cpp"
  rm -f "$MIRROR/Lib/TLibEncoder/TEncSearch.cpp"
  python "$ROOT/hm/direct/patch_hm.py" "$SRC/Lib/TLibCommon" "$MIRROR/Lib/TLibCommon" "$SRC/Lib/TLibEncoder" "$MIRROR/Lib/TLibEncoder" || exit 1
  SRC="$MIRROR"
  FLAGS="-O3 -std=c++11 -DMSYS_LINUX -w -include cmath -I$ROOT/hm/direct -I$ROOT/include -I$SRC/Lib -I$SRC/Lib/TLibCommon -I$COMMON"
  LIBSRCS=$(ls $SRC/Lib/TLibCommon/*.cpp $SRC/Lib/TLibVideoIO/*.cpp $SRC/Lib/TLibEncoder/*.cpp $SRC/Lib/TLibDecoder/*.cpp $SRC/Lib/TAppCommon/*.cpp)
  LIBSRCS="$LIBSRCS $COMMON/extraction_context.cpp $COMMON/tools.cpp $COMMON/visualization_debugging.cpp $ROOT/hm/direct/pnn_hm_direct.cpp"
  ENCSRCS=$(ls $SRC/App/TAppEncoder/*.cpp)
  DECSRCS=$(ls $SRC/App/TAppDecoder/*.cpp)
  export FLAGS OBJ
  printf '%s\n' $LIBSRCS | xargs -P "$JOBS" -I{} bash -c 'compile {} lib'
  printf '%s\n' $ENCSRCS | xargs -P "$JOBS" -I{} bash -c 'compile {} enc'
  printf '%s\n' $DECSRCS | xargs -P "$JOBS" -I{} bash -c 'compile {} dec'
  gcc -O3 -w -c "$SRC/Lib/libmd5/libmd5.c" -o "$OBJ/lib_libmd5.o"
  LINK="-L$PKG/csrc -lpnn_cuda -Wl,-rpath,\$ORIGIN/../../context_adaptive_neural_network_based_prediction_b200/csrc -lpthread -ldl"
  g++ -o "$OUT/TAppEncoderStatic_${V}_direct" $OBJ/enc_*.o $OBJ/lib_*.o $LINK
  g++ -o "$OUT/TAppDecoderStatic_${V}_direct" $OBJ/dec_*.o $OBJ/lib_*.o $LINK
  echo "built $OUT/TAppEncoderStatic_${V}_direct and $OUT/TAppDecoderStatic_${V}_direct (direct binding)"
done
cp "$REF/hevc/configuration/intra_main_rext.cfg" "$OUT/intra_main_rext.cfg"
