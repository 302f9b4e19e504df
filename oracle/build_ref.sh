#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds the UNMODIFIED reference engine (asu-trans-ai-lab/DLSim-MRM src/cpp) with g++ into
# oracle/_ref/dtalite_ref (git-ignored; it still travels to the GPU box with gpurun).
# No reference source is kept in the repository: the sources are staged into a scratch
# directory under oracle/_ref/, compiled there and the staging copy is deleted afterwards.
#
# Why not the reference's own build: CMakeLists.txt omits DTA_geometry.cpp and its
# find_package(OpenMP) fails on OpenMP_C here; the authoritative build is the MSVC project.
# Three mechanical, non-arithmetic adaptations for g++/SysV:
#   1. -include oracle/ref_shim.h : mixed-type max/min overloads (MSVC <windows.h> macros)
#   2. an empty build_config.h (config.h includes it; build_config.h.in is empty)
#   3. output.h passes the double number_of_lanes to a printf "%d" at four call sites
#      (harmless on Win64 varargs, crashes on SysV): cast to int in the staging copy.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${DTL_REFERENCE_DIR:-/root/reference}/src/cpp"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
    echo "build_ref.sh: reference sources not found at $REF (expected on the GPU box); keeping prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
STAGE="$(mktemp -d "$OUT/stage.XXXXXX")"
trap 'rm -rf "$STAGE"' EXIT
for f in DTA.h DTA_geometry.cpp DTA_geometry.h ODME.h VDF.h assignment.h config.h input.h main_api.cpp \
         output.h routing.h scenario_API.h shared_code.h simulation.cpp teestream.h trip_generation.h utils.cpp utils.h; do
    cp "$REF/$f" "$STAGE/$f"
done
: > "$STAGE/build_config.h"
sed -i -E '393s/(g_link_vector\[i\]\.number_of_lanes)/(int)\1/; 2087s/(g_link_vector\[i\]\.number_of_lanes)/(int)\1/; 3451s/(g_link_vector\[i\]\.number_of_lanes)/(int)\1/; 3530s/(g_link_vector\[i\]\.number_of_lanes)/(int)\1/' "$STAGE/output.h"
if [ "$(grep -c '(int)g_link_vector\[i\]\.number_of_lanes' "$STAGE/output.h")" != "4" ]; then
    echo "build_ref.sh: printf patch did not apply to 4 sites (reference changed?)" >&2
    exit 1
fi
CXXFLAGS="-std=c++11 -O2 -fopenmp -w -I$STAGE -include $HERE/ref_shim.h"
g++ $CXXFLAGS -c "$STAGE/utils.cpp" -o "$STAGE/utils.o" &
g++ $CXXFLAGS -c "$STAGE/simulation.cpp" -o "$STAGE/simulation.o" &
g++ $CXXFLAGS -c "$STAGE/DTA_geometry.cpp" -o "$STAGE/DTA_geometry.o" &
g++ $CXXFLAGS -c "$HERE/ref_harness.cpp" -o "$STAGE/ref_harness.o" &
wait
g++ -fopenmp -o "$OUT/dtalite_ref" "$STAGE/ref_harness.o" "$STAGE/utils.o" "$STAGE/simulation.o" "$STAGE/DTA_geometry.o"
echo "built $OUT/dtalite_ref"

# Second binary for multi-period fixtures: the committed source caps the number of demand periods with a compile-time
# constant (DTA.h:15, MAX_TIMEPERIODS = 2, and settings with 2 periods already stop); the commented-out line next to it
# (DTA.h:17) is the larger setting.  Constant-only change in the staging copy, no arithmetic touched.
sed -i -E '15s/MAX_TIMEPERIODS = 2;/MAX_TIMEPERIODS = 4;/' "$STAGE/DTA.h"
grep -q 'MAX_TIMEPERIODS = 4;' "$STAGE/DTA.h" || { echo "build_ref.sh: MAX_TIMEPERIODS patch did not apply" >&2; exit 1; }
g++ $CXXFLAGS -c "$STAGE/utils.cpp" -o "$STAGE/utils.o" &
g++ $CXXFLAGS -c "$STAGE/simulation.cpp" -o "$STAGE/simulation.o" &
g++ $CXXFLAGS -c "$STAGE/DTA_geometry.cpp" -o "$STAGE/DTA_geometry.o" &
g++ $CXXFLAGS -c "$HERE/ref_harness.cpp" -o "$STAGE/ref_harness.o" &
wait
g++ -fopenmp -o "$OUT/dtalite_ref_tp4" "$STAGE/ref_harness.o" "$STAGE/utils.o" "$STAGE/simulation.o" "$STAGE/DTA_geometry.o"
echo "built $OUT/dtalite_ref_tp4"
