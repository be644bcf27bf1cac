O=gpurun_out/r2c; mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > $O/tests.log 2>&1
b() { cfg=$1; pairs=$2; tag=$3; shift 3; timeout 300 python bench.py --config $cfg --pairs $pairs --steps 5 --warmup 3 --no-cpu-baseline "$@" > $O/bench_${cfg}_${tag}.json 2> $O/bench_${cfg}_${tag}.err; }
b C4 2000000 link
b C3 2000000 link
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4.csv python tools/dbg_map.py - C4 1000000 > $O/ncu_launch.log 2>&1
ls $O
