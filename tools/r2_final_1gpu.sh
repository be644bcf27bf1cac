# round-2 evidence at HEAD on one B200: tests, bench lines, launch list, full ncu captures, sanitizer, CLI end to end
O=gpurun_out/r2final; mkdir -p $O
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > $O/tests.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench_c4_1gpu.json 2> $O/bench_c4_1gpu.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_c4_reference_arm.json 2> $O/bench_c4_reference_arm.err
timeout 600 python bench.py --config C3 --no-cpu-baseline > $O/bench_c3_1gpu.json 2> $O/bench_c3_1gpu.err
timeout 600 python bench.py --config C2 --no-cpu-baseline --steps 20 > $O/bench_c2_1gpu.json 2> $O/bench_c2_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/launches_bench.log 2>&1
KRE='regex:k_(scan_rows|tile_sum|tile_fix|scan_redo|memo|walk|map_fast|map_windows|intern_slots|pair_agg|comb_weigh|comb_emit|list_weigh|list_emit|wkey_hist|wkey_scatter|bucket_count)'
timeout 1500 ncu --set full --clock-control none --import-source on -k "$KRE" -s 78 -c 26 -f -o /tmp/prof_c4_block python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_c4_block.log 2>&1
# the report itself is too large to bring back: summarise it here
python tools/ncu_table.py /tmp/prof_c4_block.ncu-rep > $O/ncu_kernels.txt 2>&1
for k in k_scan_rows k_memo k_walk k_map_fast k_pair_agg k_comb_emit k_comb_weigh k_wkey_scatter k_bucket_count; do
  (echo "# source-level hot lines: $k"; python tools/prof_summary.py /tmp/prof_c4_block.ncu-rep $k 30 2>&1 | tail -32) > $O/ncu_hot_$k.txt
done
(timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -12) > $O/sanitizer_memcheck.txt 2>&1
(timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | tail -12) > $O/sanitizer_racecheck.txt 2>&1
timeout 900 python tools/cli_e2e.py C4 10000000 1 > $O/cli_e2e_c4_10M.json 2> $O/cli_e2e.err
ls -la $O
