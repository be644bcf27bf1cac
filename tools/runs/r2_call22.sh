O=gpurun_out/r2z; mkdir -p $O
b() { tag=$1; shift; timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err; }
b base
for v in mf1 mf6 mf8 mm1 mm6 mm8 wk8 sm8x5 sm5x8; do VSPE_LIB_PATH=$PWD/build/variants/libvspe_$v.so b $v; done
ls $O
