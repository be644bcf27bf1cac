O=gpurun_out/r2h; mkdir -p $O
(timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scan_mode4 or chunked or edge_shapes or noisy or per_read" 2>&1 | tail -30) > $O/tests.log 2>&1
b() { cfg=$1; pairs=$2; tag=$3; shift 3; timeout 300 python bench.py --config $cfg --pairs $pairs --steps 5 --warmup 3 --no-cpu-baseline "$@" > $O/bench_${cfg}_${tag}.json 2> $O/bench_${cfg}_${tag}.err; }
b C4 2000000 fused --opt scan_mode=4
b C3 2000000 fused --opt scan_mode=4
b C2 1000000 fused --opt scan_mode=4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_scan_rows|k_walk" -s 4 -c 2 -f -o $O/prof_scanmap python tools/dbg_map.py scan_mode=4 C4 1000000 > $O/ncu_scanmap.log 2>&1
ls $O
