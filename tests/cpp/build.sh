#!/bin/bash
# builds the C++ API test against the in-tree library (compile + link check on CPU; run on the GPU box)
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$(dirname "$HERE")")"
mkdir -p "$HERE/bin"
g++ -std=c++17 -O2 -I"$ROOT/include" "$HERE/test_knowhere_api.cc" -o "$HERE/bin/test_knowhere_api" \
    -L"$ROOT/knowhere_b200" -l:libknowhere_b200.so -Wl,-rpath,"$ROOT/knowhere_b200" -Wl,-rpath,'$ORIGIN/../../../knowhere_b200'
