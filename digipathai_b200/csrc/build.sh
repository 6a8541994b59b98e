#!/usr/bin/env bash
# Builds libdigipath_b200.so in-tree for sm_100a (the only target). Usage: build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC -shared -cudart static "$@" \
     -o ../libdigipath_b200.so runtime.cu
