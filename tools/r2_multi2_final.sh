# 2-GPU evidence at HEAD: every GPU test (the N > 1 ones run here), C4 and C5 bench lines at 2 GPUs
N=2; O=gpurun_out/r2m2f; mkdir -p $O
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > $O/tests.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_c4_n$N.json 2> $O/bench_c4_n$N.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --config C5 --steps 3 --warmup 3 > $O/bench_c5_n$N.json 2> $O/bench_c5_n$N.err
ls $O
