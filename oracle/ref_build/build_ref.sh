#!/usr/bin/env bash
# Builds the UNMODIFIED-in-behaviour reference (Henning1/resql) as the parity oracle.
# TEST INFRASTRUCTURE ONLY. Reads sources from /root/reference (read-only), works in a scratch
# copy under /tmp, and writes ONLY binaries into oracle/_ref/ (git-ignored, travels via gpurun).
# No reference source is copied into the repository. The reference's own Makefile/cmake are not
# used: asmjit is compiled file by file with g++, the parser with the vendored lemon.
#
# Adaptations (all mechanical, documented in DESIGN.md "Oracle"):
#  1. lemon from lib/lemon/lemon.c generates parser.c/parser.h from src/parser/parser.y
#  2. flex is absent -> oracle/ref_build/lexer_hand.c implements src/parser/lexer.y
#  3. nasm is absent -> JitConfig::emitMachineCode defaults to true (asmjit path)
#  4. asmjit path sign-extends INT->BIGINT as 16 bit (movsx r64,r32 is not encodable); the default
#     nasm path assembles it as movsxd. Patch the single call site to movsxd so the oracle has the
#     reference's DEFAULT semantics (ExpressionsJitFlounder.h:818-824, ValuesJitFlounder.h:82-93)
#  5. g++ 13 rejects `T<N>(const T<N>&)` ctor declarations in C++20: 6 asmjit headers, macro arg only
#  6. src/resql.cpp needs readline -> oracle/ref_build/ref_driver.cpp drives executeStatement()
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../_ref"
REF="${RESQL_REFERENCE:-/root/reference}"
B="${RESQL_REF_BUILD_DIR:-/tmp/resql_ref_build}"
JOBS="${JOBS:-8}"
if [ ! -d "$REF/src" ]; then echo "reference not found at $REF" >&2; exit 3; fi
mkdir -p "$OUT"
rm -rf "$B"; mkdir -p "$B"
cp -r "$REF/src" "$REF/lib" "$REF/test" "$B/"
chmod -R u+w "$B"
cd "$B"
# 1. parser
cc -O1 -w -o lemon2 lib/lemon/lemon.c
./lemon2 -Tlib/lemon/lempar.c src/parser/parser.y -d. >/dev/null 2>&1 || ./lemon2 src/parser/parser.y -d. || true
test -f parser.c && test -f parser.h
# 5. asmjit headers for g++ C++20
sed -i -E 's/ASMJIT_NONCOPYABLE\((ZoneTmp|StringTmp|ZoneHash|ZoneVector|ZoneStack|RALiveSpans)<[A-Za-z]+>\)/ASMJIT_NONCOPYABLE(\1)/' \
   lib/asmjit/src/asmjit/core/{zone.h,string.h,zonehash.h,zonevector.h,zonestack.h,radefs_p.h}
# 3./4.
sed -i 's/bool emitMachineCode = false;/bool emitMachineCode = true;/' src/JitContextFlounder.h
sed -i 's/ctx.yield ( movsx ( res, child ) );/ctx.yield ( movsxd ( res, child ) );/' src/ExpressionsJitFlounder.h
grep -q 'movsxd ( res, child )' src/ExpressionsJitFlounder.h
# asmjit static library, file by file
mkdir -p aj
ls lib/asmjit/src/asmjit/core/*.cpp lib/asmjit/src/asmjit/x86/*.cpp | \
  xargs -P "$JOBS" -I{} sh -c 'g++ -O2 -std=c++17 -fPIC -w -DASMJIT_STATIC -DNDEBUG -Ilib/asmjit/src -c {} -o aj/$(echo {} | tr "/" "_").o'
ar rcs libasmjit.a aj/*.o
# 2. lexer
gcc -c -O2 -w -I. -Isrc "$HERE/lexer_hand.c" -o lexer_hand.o
CXX="g++ -O3 -DNDEBUG -std=c++20 -pthread -fPIC -w -DASMJIT_STATIC -Ilib/cereal/include -Ilib/cxxopts -Ilib/asmjit/src -Isrc -I."
$CXX "$HERE/ref_driver.cpp" lexer_hand.o libasmjit.a -Wl,--export-dynamic -lrt -o "$OUT/resql-oracle" &
$CXX -Itest "$HERE/ref_tests.cpp" lexer_hand.o libasmjit.a -Wl,--export-dynamic -lrt -o "$OUT/resql-reftests" &
wait
ls -la "$OUT"
