O=gpurun_out/r2l; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > $O/tests.log 2>&1
(VSPE_TEST_C5=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c5_stress" 2>&1 | tail -15) > $O/tests_c5.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
timeout 900 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
timeout 300 python bench.py --config C3 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
ls $O
