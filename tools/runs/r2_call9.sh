O=gpurun_out/r2i; mkdir -p $O
(timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scan_mode4 or chunked or edge_shapes or noisy or per_read" 2>&1 | tail -30) > $O/tests.log 2>&1
for v in A B C D E; do
  if [ $v = A ]; then unset VSPE_LIB_PATH; else export VSPE_LIB_PATH=$PWD/vstrains_b200/libvspe_$v.so; fi
  for cfg in C4 C2; do
    timeout 300 python bench.py --config $cfg --pairs 2000000 --steps 5 --warmup 3 --no-cpu-baseline --opt scan_mode=4 > $O/bench_${cfg}_$v.json 2> $O/bench_${cfg}_$v.err
  done
done
unset VSPE_LIB_PATH
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_scan_rows|k_walk" -s 4 -c 2 -f -o $O/prof_scanmap python tools/dbg_map.py scan_mode=4 C4 1000000 > $O/ncu_scanmap.log 2>&1
ls $O
