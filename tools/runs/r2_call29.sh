O=gpurun_out/r2ah; mkdir -p $O
b() { tag=$1; shift; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err; }
b c4
b c2 --config C2
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5) > $O/tests.log 2>&1
ls $O
