# round 2, call 1: what never ran (two_err), candidate timings, ncu evidence of the round-1 kernels at HEAD
O=gpurun_out/r2a; mkdir -p $O
(nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g; lscpu | head -20; nvidia-smi topo -m) > $O/box.txt 2>&1
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_err" 2>&1 | tail -15) > $O/tests_two_err.log 2>&1
b() { cfg=$1; pairs=$2; tag=$3; shift 3; timeout 300 python bench.py --config $cfg --pairs $pairs --steps 5 --warmup 3 --no-cpu-baseline "$@" > $O/bench_${cfg}_${tag}.json 2> $O/bench_${cfg}_${tag}.err; }
b C4 2000000 default
b C4 2000000 flat --opt count_flat=1
b C4 2000000 twoerr --opt two_err=1
b C3 2000000 default
b C3 2000000 both --opt count_flat=1 --opt two_err=1
b C2 1000000 both --opt count_flat=1 --opt two_err=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4.csv python tools/dbg_map.py - C4 1000000 > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_map_first|k_map_second|k_map_fast|k_pair_count|k_pair_emit|k_bucket_hist|k_scan_pack' -s 26 -c 13 -o $O/prof_c4 python tools/dbg_map.py - C4 1000000 > $O/ncu_full.log 2>&1
ls -la $O
