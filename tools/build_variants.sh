# kernel-parameter experiments: one library per variant under build/variants/ (selected with VSPE_LIB_PATH)
set -e
cd "$(dirname "$0")/../vstrains_b200/csrc"
mkdir -p ../../build/variants
build() { name=$1; shift; make -s clean; make -s -j EXTRA="$*" OUT=../../build/variants/libvspe_$name.so; }
build mf6 -DVSPE_MF_MINB=6
build mf8 -DVSPE_MF_MINB=8
build mm6 -DVSPE_MM_MINB=6
build mm8 -DVSPE_MM_MINB=8
build wk8 -DVSPE_WK_MINB=8
build sm8x5 -DVSPE_SM_WARPS=8 -DVSPE_SM_MINB=5
build sm5x8 -DVSPE_SM_WARPS=5 -DVSPE_SM_MINB=8
make -s clean; make -s -j
