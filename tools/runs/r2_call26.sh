O=gpurun_out/r2ad; mkdir -p $O
b() { tag=$1; shift; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err; }
for v in base scan link both both_spread1 both_ldcs; do VSPE_LIB_PATH=$PWD/build/variants/libvspe_$v.so b c4_$v; done
for v in base both; do VSPE_LIB_PATH=$PWD/build/variants/libvspe_$v.so b c3_$v --config C3; done
(VSPE_LIB_PATH=$PWD/build/variants/libvspe_both.so timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5) > $O/tests_both.log 2>&1
ls $O
