O=gpurun_out/r2k; mkdir -p $O
(VSPE_TEST_C5=1 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c5_stress" 2>&1 | tail -15) > $O/tests_c5.log 2>&1
timeout 900 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_c5_n1.json 2> $O/bench_c5_n1.err
ls $O
