#!/usr/bin/env bash
# Builds the reference's own `calc` and `scalc` (and `view`, as a .two->text dumper, and `sort`) from the
# sources where they lie under /root/reference into oracle/_ref/ (git-ignored,
# shipped to the GPU box by gpurun). TEST INFRASTRUCTURE ONLY: the product never
# links or executes anything built here. Recipe mirrors the reference makefile
# flags (makefile:104,123: -std=c++0x -O3 -msse4.2) with three workarounds:
#   * shim/zstd.h + shim/zstd_errors.h: declaration-only headers, linked against
#     the image's runtime libzstd.so.1 (no zstd dev headers in this image);
#   * forced standard includes the reference relies on transitively (GCC 13);
#   * stub_common.h replaces lib/tomahawk.cpp (needs htslib, absent here).
# `import` (htslib) is NOT built; synthetic .twk files are written by
# oracle/twk_format.py following the reference's on-disk layout.
set -euo pipefail
REF=${TWK_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/lib/ld" ]; then
  echo "[oracle] $REF not present; keeping prebuilt $OUT (if any)"; exit 0
fi
mkdir -p "$OUT/obj"
FORCE="-include cstdint -include cstring -include cassert -include limits -include algorithm -include bitset -include cmath"
CXXFLAGS="-std=c++0x -O3 -msse4.2 -w $FORCE -I$REF/lib -I$REF/include -I$HERE/shim -I$HERE -DVERSION=\"oracle\" -pthread"
SRCS="lib/buffer.cpp lib/core.cpp lib/fisher_math.cpp lib/header.cpp lib/index.cpp lib/intervals.cpp
      lib/twk_reader.cpp lib/two_reader.cpp lib/two_sorter_structs.cpp lib/utility.cpp lib/zstd_codec.cpp
      lib/ld/ld.cpp lib/ld/ld_engine.cpp lib/ld/ld_structs.cpp"
OBJS=""
pids=()
for s in $SRCS; do
  o="$OUT/obj/$(echo "$s" | tr '/' '_' | sed 's/\.cpp$/.o/')"
  OBJS="$OBJS $o"
  if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ]; then
    g++ $CXXFLAGS -c "$REF/$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
g++ $CXXFLAGS "$HERE/stub_calc_main.cpp" $OBJS -l:libzstd.so.1 -o "$OUT/tomahawk_calc" 2>/dev/null
g++ $CXXFLAGS "$HERE/stub_view_main.cpp" $OBJS -l:libzstd.so.1 -o "$OUT/tomahawk_view" 2>/dev/null
g++ $CXXFLAGS "$HERE/stub_sort_main.cpp" $OBJS -l:libzstd.so.1 -o "$OUT/tomahawk_sort" 2>/dev/null
g++ $CXXFLAGS "$HERE/stub_scalc_main.cpp" $OBJS -l:libzstd.so.1 -o "$OUT/tomahawk_scalc" 2>/dev/null
g++ $CXXFLAGS "$HERE/stub_decay_main.cpp" $OBJS -l:libzstd.so.1 -o "$OUT/tomahawk_decay" 2>/dev/null
g++ $CXXFLAGS "$HERE/stub_aggregate_main.cpp" $OBJS -l:libzstd.so.1 -o "$OUT/tomahawk_aggregate" 2>/dev/null
# The same fisher_math.cpp as a tiny shared object so the C restatement's Fisher
# can be diffed against the reference's kt_fisher_exact directly (ctypes).
g++ -O3 -msse4.2 -w -shared -fPIC -I$REF/lib "$REF/lib/fisher_math.cpp" -o "$OUT/libref_fisher.so"
echo "[oracle] built $OUT/tomahawk_calc $OUT/tomahawk_scalc $OUT/tomahawk_decay $OUT/tomahawk_aggregate $OUT/tomahawk_view $OUT/tomahawk_sort $OUT/libref_fisher.so"
