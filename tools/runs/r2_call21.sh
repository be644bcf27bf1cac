O=gpurun_out/r2y; mkdir -p $O
b() { tag=$1; shift; timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err; }
for cfg in C4 C3; do
  b ${cfg}_s64k_f0 --config $cfg
  b ${cfg}_s16k_f0 --config $cfg --opt link_split=16384
  b ${cfg}_s16k_f1 --config $cfg --opt link_split=16384 --opt link_fold=1
  b ${cfg}_s64k_f1 --config $cfg --opt link_fold=1
  b ${cfg}_snone_f0 --config $cfg --opt link_split=1073741824
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/launches_bench.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or synthetic or noisy" 2>&1 | tail -4) > $O/tests.log 2>&1
ls $O
