mkdir -p gpurun_out
(timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/tests.log 2>&1; cat gpurun_out/tests.log
timeout 100 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err
timeout 60 python bench.py --config C3 --pairs 1000000 --steps 5 --no-cpu-baseline > gpurun_out/bench_C3b.json 2> gpurun_out/bench_C3b.err
