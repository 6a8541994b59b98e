#!/usr/bin/env bash
# Builds libdigipath_b200.so and libdigipath_ingest.so in-tree for sm_100a (the only target). Usage: build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC -shared -cudart static "$@" \
     -o ../libdigipath_b200.so runtime.cu
# ingest library (nvJPEG tile decode + scatter into the [x][y][c] raster); nvJPEG linked statically
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC -shared -cudart static "$@" \
     -o ../libdigipath_ingest.so ingest.cu -lnvjpeg_static -lculibos
