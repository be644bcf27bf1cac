mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/tests.log 2>&1; cat gpurun_out/tests.log
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/sw_default.json 2> gpurun_out/sw_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_y.csv python tools/dbg_map.py - C2 1000000 > gpurun_out/ncu_y.log 2>&1
