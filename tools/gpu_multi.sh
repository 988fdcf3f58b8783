#!/bin/bash
# N-GPU session: NCCL slab parity test + bench under torchrun. usage: gpu_multi.sh TAG NGPU
TAG=${1:-m}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q -s > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 $OUT/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 > $OUT/bench_$N.json 2> $OUT/bench_$N.err; echo "bench $N rc=$?"; cat $OUT/bench_$N.json; tail -5 $OUT/bench_$N.err
