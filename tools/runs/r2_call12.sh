O=gpurun_out/r2n; mkdir -p $O
b() { tag=$1; shift; timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err; }
b c4_ov1_cap21
b c4_ov0_cap21 --opt tier_overlap=0
b c4_ov0_cap0 --opt tier_overlap=0 --opt pair_cap_log2=0
b c4_ov0_cap23 --opt tier_overlap=0 --opt pair_cap_log2=23
b c4_ov1_cap0 --opt pair_cap_log2=0
b c3_ov0_cap21 --config C3 --opt tier_overlap=0
b c3_ov1_cap21 --config C3
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gzip" 2>&1 | tail -5) > $O/tests_gz.log 2>&1
ls $O
