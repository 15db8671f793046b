#!/bin/bash
# Builds the stand-alone GEMM probes (dev tools) into tools/_bin/ for sm_100a.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_bin
CS=ace-step-1.5-for-windows_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -lcuda -DACE_PROBE"
nvcc $FLAGS -o tools/_bin/gemm_probe tools/gemm_probe.cu $CS/runtime.cu
nvcc $FLAGS -DACE_GEMM_TIMING -o tools/_bin/gemm_timing tools/gemm_timing.cu $CS/runtime.cu
nvcc $FLAGS -DACE_GEMM_TIMING -o tools/_bin/gemm_l2_probe tools/gemm_l2_probe.cu $CS/runtime.cu
nvcc $FLAGS -o tools/_bin/mufu_probe tools/mufu_probe.cu
nvcc $FLAGS -o tools/_bin/tmem_probe tools/tmem_probe.cu $CS/runtime.cu
nvcc $FLAGS -DACE_ATTN_TIMING -o tools/_bin/attn_timing tools/attn_timing.cu $CS/runtime.cu
nvcc $FLAGS -o tools/_bin/gemm_multiwave tools/gemm_multiwave.cu $CS/runtime.cu
nvcc  -o tools/_bin/desc_rowstep_probe tools/desc_rowstep_probe.cu $CS/runtime.cu
nvcc $FLAGS -DACE_RU_TIMING -o tools/_bin/ru_timing tools/ru_timing.cu $CS/runtime.cu
