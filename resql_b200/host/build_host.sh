#!/usr/bin/env bash
# Builds resql-b200: the reference front end (parser, planner, operators - compiled from the
# reference checkout, nothing copied into this repo) + our host shim + libresql_b200.so.
# Needs the scratch tree prepared by oracle/ref_build/build_ref.sh (lemon parser, asmjit archive).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$HERE/../.."
B="${RESQL_REF_BUILD_DIR:-/tmp/resql_ref_build}"
if [ ! -f "$B/libasmjit.a" ]; then bash "$ROOT/oracle/ref_build/build_ref.sh"; fi
test -f "$ROOT/resql_b200/libresql_b200.so"
g++ -O2 -DNDEBUG -std=c++20 -pthread -fPIC -w -DASMJIT_STATIC \
    -I"$B/lib/cereal/include" -I"$B/lib/cxxopts" -I"$B/lib/asmjit/src" -I"$B/src" -I"$B" \
    -I"$ROOT/include" -I"$HERE" \
    "$HERE/resql_b200_driver.cpp" "$B/lexer_hand.o" "$B/libasmjit.a" \
    -L"$ROOT/resql_b200" -lresql_b200 -Wl,-rpath,'$ORIGIN/..' -Wl,--export-dynamic -lrt \
    -o "$HERE/resql-b200"
ls -la "$HERE/resql-b200"
# the reference's own test suites routed through the drop-in (ref_tests_gpu.cpp)
g++ -O2 -DNDEBUG -std=c++20 -pthread -fPIC -w -DASMJIT_STATIC \
    -I"$B/lib/cereal/include" -I"$B/lib/cxxopts" -I"$B/lib/asmjit/src" -I"$B/src" -I"$B" -I"$B/test" \
    -I"$ROOT/include" -I"$HERE" \
    "$HERE/ref_tests_gpu.cpp" "$B/lexer_hand.o" "$B/libasmjit.a" \
    -L"$ROOT/resql_b200" -lresql_b200 -Wl,-rpath,'$ORIGIN/..' -Wl,--export-dynamic -lrt \
    -o "$HERE/resql-reftests-gpu"
ls -la "$HERE/resql-reftests-gpu"
