#!/bin/bash
# N-GPU validation of bench.py under torchrun (weak scaling: 1M source points per GPU, NCCL all-reduce of H|b|e).
N=${1:-2}
TAG=${2:-r01_multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for n in $(seq 1 1); do :; done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "rc=$?"; tail -c 2500 $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err
echo "rc=$?"; tail -c 600 $OUT/bench_ref_n$N.json
