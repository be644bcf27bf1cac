# evidence at HEAD: GPU tests, launch list of the bench command, ncu --set full of every kernel of a step,
# compute-sanitizer, CLI end to end
O=gpurun_out/r2q; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/tests.log 2>&1
KRE='regex:k_(scan_rows|walk|map_fast|map_windows|pair_agg|comb_weigh|comb_emit|list_weigh|list_emit|wkey_hist|wkey_scatter|bucket_count)'
# launch list of the bench command (one warm-up + one timed step of the full-size C4 workload)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_c4.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/launches_bench.log 2>&1
# full captures: the kernels of one 10 M-pair block of the bench command (the second block: warm tables)
timeout 1500 ncu --set full --clock-control none --import-source on -k "$KRE" -s 30 -c 30 -f -o $O/prof_c4_block \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_c4_block.log 2>&1
# compute-sanitizer on the small end-to-end run
(timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -12) > $O/sanitizer_memcheck.txt 2>&1
(timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | tail -12) > $O/sanitizer_racecheck.txt 2>&1
# CLI end to end, C4 block (10 M pairs, 6.3 GB of FASTQ), files on disk -> pe_info / st_info on disk
timeout 900 python tools/cli_e2e.py C4 10000000 1 > $O/cli_e2e_c4_10M.json 2> $O/cli_e2e.err
ls -la $O
