#!/bin/bash
# Build the torch-free GPU self-test / probe binaries in-tree (git-ignored, but they travel to the GPU box with gpurun).
set -e
cd "$(dirname "$0")/.."
nvcc -O2 -std=c++17 -Wno-deprecated-gpu-targets -o tools/cu/tc_selftest.bin tools/cu/tc_selftest.cu -ldl
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/cu/umma_probe.bin tools/cu/umma_mn_major_probe.cu
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/cu/umma_2cta_probe.bin tools/cu/umma_2cta_probe.cu
ls -la tools/cu/*.bin
