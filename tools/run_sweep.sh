mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "synthetic or golden or full_size or accumulate" 2>&1 | tail -5) > gpurun_out/tests.log 2>&1; cat gpurun_out/tests.log
for o in "count_low_bits=7" "count_low_bits=9" "count_low_bits=11" "count_low_bits=13"; do
  tag=$(echo "$o" | tr ' =-' '___')
  timeout 300 python bench.py --no-cpu-baseline --steps 8 --opt $o > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err
done
