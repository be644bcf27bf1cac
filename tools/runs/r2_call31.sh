O=gpurun_out/r2ai; mkdir -p $O
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_c4_off.json 2> $O/bench_c4_off.err
VSPE_L2_PERSIST=1 timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_c4_on.json 2> $O/bench_c4_on.err
