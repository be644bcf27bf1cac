mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/tests.log 2>&1; cat gpurun_out/tests.log
for o in "list_spread=2" "list_spread=1" "list_spread=2 --opt chunk_mb=64" "list_spread=2 --opt chunk_mb=32" "list_spread=2 --opt chunk_mb=128"; do
  tag=$(echo "$o" | tr ' =-' '___')
  python bench.py --no-cpu-baseline --steps 8 --opt $o > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err
done
