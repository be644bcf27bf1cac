O=gpurun_out/r2d; mkdir -p $O
(timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scan_mode4 or chunked or edge_shapes or noisy or per_read" 2>&1 | tail -40) > $O/tests.log 2>&1
b() { cfg=$1; pairs=$2; tag=$3; shift 3; timeout 300 python bench.py --config $cfg --pairs $pairs --steps 5 --warmup 3 --no-cpu-baseline "$@" > $O/bench_${cfg}_${tag}.json 2> $O/bench_${cfg}_${tag}.err; }
b C4 2000000 fused --opt scan_mode=4
b C3 2000000 fused --opt scan_mode=4
b C2 1000000 fused --opt scan_mode=4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c4.csv python tools/dbg_map.py scan_mode=4 C4 1000000 > $O/ncu_launch.log 2>&1
ls $O
