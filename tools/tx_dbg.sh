#!/usr/bin/env bash
mkdir -p gpurun_out/r2x7
for d in 0 1 2 4 6 8 16 30; do
  DP_TX_DBG=$d timeout 200 python tools/tf32x3_bringup.py dense 2>&1 | grep -E "tf32x3:|dec10b|dec7b|dec9b|conv2_block6" | tr '\n' ' ' | sed "s/^/dbg $d: /"; echo
done
