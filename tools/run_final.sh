# Round-end measurement set (one gpurun call): parity tests, the bench line, the reference arm,
# other configs, the ncu launch list of the bench command.
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/tests.log 2>&1; cat gpurun_out/tests.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
for c in C3 C4; do timeout 300 python bench.py --config $c --pairs 1000000 --steps 5 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c3.csv python tools/dbg_map.py - C3 500000 > gpurun_out/ncu_c3.log 2>&1
