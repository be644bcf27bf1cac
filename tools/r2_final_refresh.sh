# refresh of the headline evidence at the final HEAD: GPU tests, smoke, default bench line, launch list
O=gpurun_out/r2final2; mkdir -p $O
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $O/tests.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench_c4_1gpu.json 2> $O/bench_c4_1gpu.err
timeout 600 python bench.py --config C3 --no-cpu-baseline > $O/bench_c3_1gpu.json 2> $O/bench_c3_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/launches_bench.log 2>&1
KRE='regex:k_(comb_weigh|comb_emit)'
timeout 600 ncu --set full --clock-control none --import-source on -k "$KRE" -s 20 -c 2 -f -o /tmp/prof_link python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_link.log 2>&1
python tools/ncu_table.py /tmp/prof_link.ncu-rep > $O/ncu_link_kernels.txt 2>&1
ls $O
