# 8-GPU evidence: N>1 parity tests, C4 bench at 8 GPUs, C5 (200k-node stress graph, sparse) at 8 GPUs
N=8; O=gpurun_out/r2m8; mkdir -p $O
(nvidia-smi topo -m; nproc; free -g) > $O/topo.txt 2>&1
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "torchrun or multi_gpu" 2>&1 | tail -15) > $O/tests.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_c4_n$N.json 2> $O/bench_c4_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --config C5 --steps 3 --warmup 3 > $O/bench_c5_n$N.json 2> $O/bench_c5_n$N.err
ls $O
