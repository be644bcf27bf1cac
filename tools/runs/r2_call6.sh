O=gpurun_out/r2e; mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_blocks or many_long or sparse or run_rank" 2>&1 | tail -25) > $O/tests.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err
ls $O
