#!/usr/bin/env bash
# Builds the C++ host over the C ABI: libracer_host.so (scene/config loading, flattening, host BVH, the
# Renderer mirror, PNG output) and the headless driver racer_render, both in-tree next to libracer_cuda.so.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
PKG="${HERE}/.."
CXX="${CXX:-g++}"
FLAGS="-O2 -std=c++17 -fPIC -Wall -Wextra -Wno-missing-field-initializers"
"${CXX}" ${FLAGS} -shared -o "${PKG}/libracer_host.so" "${HERE}/racer_host.cpp" -L"${PKG}" -lracer_cuda -lz -Wl,-rpath,'$ORIGIN'
"${CXX}" ${FLAGS} -o "${PKG}/racer_render" "${HERE}/main.cpp" -L"${PKG}" -lracer_host -lracer_cuda -Wl,-rpath,'$ORIGIN'
echo "built ${PKG}/libracer_host.so ${PKG}/racer_render"
