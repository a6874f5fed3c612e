#!/bin/bash
# tools/umma_probe_sweep.sh -- build tools/cu/umma_mn_major_probe.cu and try descriptor candidates, one process each
# (a malformed descriptor may poison the CUDA context).  Usage on the GPU box:  bash tools/umma_probe_sweep.sh | tee gpurun_out/umma_probe.txt
set -u
BIN=/tmp/umma_probe
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o $BIN "$(dirname "$0")/cu/umma_mn_major_probe.cu" || exit 1
echo "# control: both operands K-major (the validated encoding) -- must MATCH or the harness is broken"
timeout 30 $BIN 0 0 16 1024 32
echo "# expected encoding first (CUTLASS convention); the sweep only runs if it is wrong"
if timeout 30 $BIN 1 1 8192 1024 2048; then echo "# expected encoding confirmed"; exit 0; fi
echo "# MN-major candidates"
for ab in "1 1"; do
  for lbo in 8192 1024 128 16; do
    for sbo in 1024 8192 128; do
      for kstep in 2048 32 4096; do
        timeout 30 $BIN $ab $lbo $sbo $kstep | grep -E "MATCH|ERROR" 
      done
    done
  done
done
echo "# done (lines above list only matching or faulting candidates)"
