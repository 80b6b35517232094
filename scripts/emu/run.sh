#!/bin/bash
# DEVELOPMENT TOOL (see include/cuda_runtime.h).  Dry run of the gpu-marked tests in a GPU-less container:
# copies the repository to $EMU_ROOT (default /tmp/chimp_emu/repo), builds the engine there as a host model in
# place of libchimp_b200.so and runs pytest on the copy with the torch device shim.  The working tree is not touched.
#   scripts/emu/run.sh [pytest arguments, default: tests -m gpu -x -q]
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$(cd "$HERE/../.." && pwd)
ROOT=${EMU_ROOT:-/tmp/chimp_emu/repo}
OBJ=${EMU_OBJ:-/tmp/chimp_emu/obj}
rm -rf "$ROOT" && mkdir -p "$ROOT" "$OBJ"
tar -C "$SRC" --exclude ./.git --exclude ./gpurun_out --exclude __pycache__ --exclude .pytest_cache \
    --exclude ./badchimp-cpp_b200/libchimp_b200.so --exclude ./badchimp-cpp_b200/build -cf - . | tar -C "$ROOT" -xf -
PKG="$ROOT/badchimp-cpp_b200"
python "$HERE/prep.py" "$PKG/csrc" "$PKG/csrc_emu"
# EMU_SANITIZE=1: AddressSanitizer + UBSan build of the model (out-of-bounds reads of the kernels and of the host code)
SAN=""
if [ -n "$EMU_SANITIZE" ]; then SAN="-fsanitize=address,undefined -fno-omit-frame-pointer -fno-sanitize-recover=undefined"; fi
FLAGS="$SAN -O2 -g -std=c++17 -fPIC -ffp-contract=off -fno-strict-aliasing -I$HERE/include -pthread -w -include cuda_runtime.h"
# object cache keyed by the content of the translated sources and of the model
KEY=$(echo "$SAN" | cat - "$PKG"/csrc_emu/* "$HERE"/include/cuda_runtime.h "$HERE"/emu_runtime.cpp "$SRC"/include/chimp_b200.h | md5sum | cut -c1-16)
if [ ! -f "$OBJ/$KEY.so" ]; then
  rm -f "$OBJ"/*.o
  g++ $FLAGS -c "$PKG/csrc_emu/engine.cpp" -o "$OBJ/engine.o" & p1=$!
  g++ $FLAGS -c "$PKG/csrc_emu/kernels.cpp" -o "$OBJ/kernels.o" & p2=$!
  g++ $FLAGS -c "$HERE/emu_runtime.cpp" -o "$OBJ/emu_runtime.o" & p3=$!
  wait $p1; wait $p2; wait $p3
  # -Bsymbolic: the library's cuda* calls bind to its own host model even when torch has the real runtime loaded
  g++ $SAN -shared -Wl,-Bsymbolic -o "$OBJ/$KEY.so" "$OBJ/engine.o" "$OBJ/kernels.o" "$OBJ/emu_runtime.o" -pthread
fi
cp "$OBJ/$KEY.so" "$PKG/libchimp_b200.so"
# host apps link "-lcudart": a linker script under that name hands them the model's runtime (one copy of its state);
# binaries built against the real library are dropped from the copy, the tests rebuild them
mkdir -p "$ROOT/.emu_lib" && echo "INPUT ( $PKG/libchimp_b200.so )" > "$ROOT/.emu_lib/libcudart.so"
for app in std_case dump_tables std_one_phase twophase vtk_write voxel_case cpu_loop; do rm -f "$PKG/host/apps/$app"; done
sed -i "s|-L/usr/local/cuda/lib64|-L$ROOT/.emu_lib|g; s|-I/usr/local/cuda/include|-I$HERE/include|g" "$ROOT"/tests/*.py "$ROOT"/__graft_entry__.py "$ROOT"/oracle/Makefile
cd "$ROOT"
export CHIMP_EMU=1 PYTHONPATH="$HERE:$PYTHONPATH"
if [ -n "$EMU_SANITIZE" ]; then
  export LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0
fi
if [ $# -eq 0 ]; then set -- tests -m gpu -x -q; fi
exec python -m pytest -p emu_plugin "$@"
