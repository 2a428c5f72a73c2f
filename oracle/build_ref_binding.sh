#!/bin/bash
# TEST INFRASTRUCTURE.  Builds the reference's own pybind11 module `alphazero_cpp` (L1+L2:
# rules engine + adapter + Node + libtorch encoder) from the sources where they lie under
# /root/reference, for one geometry, into oracle/_ref/binding_R<R>/alphazero_cpp.so.
# Only used in the build container by tests/golden/make_golden.py to dump golden fixtures
# (encoder planes, legal masks, MCTS node statistics).  The reference's setup.py is
# MSVC-only (setup.py:30-31), hence the hand build.
# usage: oracle/build_ref_binding.sh R IA
set -euo pipefail
R=${1:-14}; IA=${2:-3}
REF=/root/reference/src/cpp
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref/binding_R$R
mkdir -p "$OUT"
PY=$(python -c 'import sysconfig; print(sysconfig.get_paths()["include"])')
TORCH=$(python -c 'import torch, os; print(os.path.dirname(torch.__file__))')
PYB=$(python -c 'import pybind11; print(pybind11.get_include())')
FLAGS="-std=c++17 -O2 -fPIC -fpermissive -w -I$REF -I$PY -I$TORCH/include -I$TORCH/include/torch/csrc/api/include -I$PYB -DTORCH_EXTENSION_NAME=alphazero_cpp -D_GLIBCXX_USE_CXX11_ABI=1"
geom_sed() {
  sed -e "s/constexpr int rows_ = 8;/constexpr int rows_ = $R;/" \
      -e "s/constexpr int cols_ = 8;/constexpr int cols_ = $R;/" \
      -e "s/constexpr int invalid_area = 2;/constexpr int invalid_area = $IA;/"
}
build_tu() { # src obj extra
  set -o pipefail
  g++ $FLAGS $3 -E "$1" | geom_sed | g++ $FLAGS $3 -x c++-cpp-output -c -o "$2" - ; }
build_tu $REF/engine/board.cpp "$OUT/engine.o" -fkeep-inline-functions &
build_tu $REF/board.cpp "$OUT/board.o" "" &
build_tu $REF/move.cpp "$OUT/move.o" "" &
build_tu $REF/node.cpp "$OUT/node.o" "" &
build_tu $REF/wrapper.cpp "$OUT/wrapper.o" "" &
for j in $(jobs -p); do wait "$j"; done
g++ -shared -o "$OUT/alphazero_cpp.so" "$OUT"/engine.o "$OUT"/board.o "$OUT"/move.o "$OUT"/node.o "$OUT"/wrapper.o \
  -L"$TORCH/lib" -ltorch -ltorch_cpu -lc10 -ltorch_python -Wl,-rpath,"$TORCH/lib"
rm -f "$OUT"/*.o
printf 'def profile(f):\n    return f\n' > "$OUT/line_profiler_pycharm.py"
echo "built $OUT/alphazero_cpp.so"
