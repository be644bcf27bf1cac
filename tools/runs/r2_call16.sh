O=gpurun_out/r2t; mkdir -p $O
(timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8) > $O/tests.log 2>&1
b() { tag=$1; shift; timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err; }
b c4
KRE='regex:k_(scan_rows|walk|tile_fix|tile_sum|scan_redo)'
timeout 900 ncu --set full --clock-control none --import-source on -k "$KRE" -s 10 -c 5 -f -o $O/prof_scan_walk python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu.log 2>&1
ls $O
