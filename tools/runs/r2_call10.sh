O=gpurun_out/r2j; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > $O/tests.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
timeout 300 python bench.py --config C3 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c4.csv python tools/dbg_map.py - C4 4000000 > $O/ncu_launch.log 2>&1
ls $O
