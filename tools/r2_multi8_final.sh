# 8-GPU bench line at the final HEAD (weak scaling: every rank its own 50 M-pair step)
N=8; O=gpurun_out/r2m8f; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_c4_n$N.json 2> $O/bench_c4_n$N.err
ls $O
