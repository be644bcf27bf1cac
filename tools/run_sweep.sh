mkdir -p gpurun_out
for o in "second_spread=1" "second_spread=4" "second_spread=8" "second_spread=16" "second_spread=8 --opt list_spread=8" "second_spread=8 --opt list_spread=16" "second_spread=8 --opt list_spread=32"; do
  tag=$(echo "$o" | tr ' =-' '___')
  python bench.py --no-cpu-baseline --steps 8 --opt $o > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err
done
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "noisy or synthetic" 2>&1 | tail -5) > gpurun_out/tests.log 2>&1; cat gpurun_out/tests.log
