O=gpurun_out/r2ac; mkdir -p $O
b() { tag=$1; shift; timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err; }
b c4
b c3 --config C3
b c2 --config C2
(timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6) > $O/tests.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/launches_bench.log 2>&1
ls $O
