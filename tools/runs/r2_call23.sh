O=gpurun_out/r2aa; mkdir -p $O
timeout 900 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_c5.json 2> $O/bench_c5.err
(VSPE_TEST_C5=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c5_stress or sparse" 2>&1 | tail -6) > $O/tests_c5.log 2>&1
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_c4.json 2> $O/bench_c4.err
ls $O
