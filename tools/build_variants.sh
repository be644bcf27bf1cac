#!/bin/bash
# Kernel-parameter experiments: builds one library per (name, nvcc flags) pair under build/variants/ and restores the
# default library.  A variant is selected at run time with VSPE_LIB_PATH=build/variants/libvspe_<name>.so (the
# directory is git-ignored but travels with gpurun).  Tunables: VSPE_SM_WARPS, VSPE_SM_ITERS, VSPE_SM_MINB (k_scan_rows),
# VSPE_MM_MINB (k_memo), VSPE_WK_MINB (k_walk), VSPE_MF_MINB (k_map_fast).
# usage: tools/build_variants.sh mm6 "-DVSPE_MM_MINB=6" sm8x5 "-DVSPE_SM_WARPS=8 -DVSPE_SM_MINB=5"
set -e
cd "$(dirname "$0")/../vstrains_b200/csrc"
mkdir -p ../../build/variants
while [ $# -ge 2 ]; do
  make -s clean
  make -s -j EXTRA="$2" OUT=../../build/variants/libvspe_$1.so
  shift 2
done
make -s clean
make -s -j
