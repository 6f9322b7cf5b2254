#!/bin/sh
# TEST INFRASTRUCTURE ONLY: compiles the product kernel sources with g++ against the SIMT emulator
# (tests/emu/simt_emu.h) into tests/emu/libsccg_b200_emu.so so kernel logic can be unit-tested
# without a GPU.  The product library is built by __graft_entry__.build() with nvcc for sm_100a.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
CXX_BIN=/usr/bin/g++
[ -x "$CXX_BIN" ] || CXX_BIN=g++
OUT="$HERE/libsccg_b200_emu.so"
SRC="$ROOT/sccg-genome-compression_b200/csrc"
NEWEST=$(ls -t "$SRC"/* "$HERE"/simt_emu.* "$ROOT/include/sccg.h" | head -1)
if [ -f "$OUT" ] && [ "$OUT" -nt "$NEWEST" ]; then exit 0; fi
# (built next to the target and renamed: test workers that start at the same time never load a half-written library)
TMP="$OUT.tmp.$$"
"$CXX_BIN" -O2 -g -std=c++17 -fPIC -shared -DSCCG_EMU $SCCG_EMU_EXTRA -Wall -Wno-unused-function -Wno-unused-parameter -Wno-unknown-pragmas \
    -I"$HERE" -I"$SRC" -x c++ "$SRC/sccg_b200.cu" -x c++ "$HERE/simt_emu.cpp" -o "$TMP" -lpthread
mv -f "$TMP" "$OUT"
echo "built $OUT"
