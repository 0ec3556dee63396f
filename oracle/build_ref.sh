#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Compiles the UNMODIFIED v1 reference renderer (the only generation of the reference that
# compiles and renders) from the sources where they lie under
# $REF (/root/reference/old), plus oracle/ref_harness.cpp, into oracle/_ref/:
#   libref_v1_strict.so  -O2, no fast-math  -> parity second opinion (double precision)
#   libref_v1_fast.so    the reference's own flags (old/setup copy.py:29-47) -> CPU timing
# Nothing from the reference is copied: the two ".h" files the reference sources include
# by their un-suffixed names are generated one-line forwarding headers.
# /root/reference does not exist on the GPU box; there the prebuilt .so files are used.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference/old}"
OUT="$HERE/_ref"
if [ ! -f "$REF/raytracer_core copy.cpp" ]; then
    echo "build_ref.sh: reference sources not found under $REF (fine on the GPU box)" >&2
    exit 3
fi
mkdir -p "$OUT/inc"
printf '#pragma once\n#include "%s/raytracer_core copy.h"\n' "$REF" > "$OUT/inc/raytracer_core.h"
printf '#pragma once\n#include "%s/bvh copy.h"\n' "$REF" > "$OUT/inc/bvh.h"
SRCS=("$HERE/ref_harness.cpp" "$REF/raytracer_core copy.cpp" "$REF/bvh copy.cpp")
COMMON=(-std=c++17 -fPIC -shared -fopenmp -I"$OUT/inc" -w)
# strict: IEEE double arithmetic, used to check hit ids / distances / images
g++ -O2 "${COMMON[@]}" "${SRCS[@]}" -o "$OUT/libref_v1_strict.so"
# fast: the reference's own optimisation flags.  -march=native/-mtune=native are replaced by
# x86-64-v3 (AVX2+FMA, what its -mavx -mfma asks for) so the binary also runs on the GPU
# box's host CPU, which need not be the CPU it was compiled on.
g++ -O3 -march=x86-64-v3 -ffast-math -funroll-loops -ftree-vectorize -fno-trapping-math \
    -fopenmp-simd -fomit-frame-pointer -msse4.2 -mavx -mfma \
    "${COMMON[@]}" "${SRCS[@]}" -o "$OUT/libref_v1_fast.so"
echo "built $OUT/libref_v1_strict.so $OUT/libref_v1_fast.so"
