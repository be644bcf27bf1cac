O=gpurun_out/r2ae; mkdir -p $O
b() { tag=$1; shift; timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err; }
b chunk256
b chunk128 --opt chunk_mb=128
b chunk64 --opt chunk_mb=64
b chunk512 --opt chunk_mb=512
ls $O
