O=gpurun_out/r2af; mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "caller_buffers or device_resident or golden" 2>&1 | tail -4) > $O/tests.log 2>&1
timeout 900 python bench.py > $O/bench_c4_1gpu.json 2> $O/bench_c4_1gpu.err
ls $O
