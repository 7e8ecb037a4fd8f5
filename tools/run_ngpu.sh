#!/bin/bash
# bench.py on N GPUs of one box, launched the way the driver does it: bash tools/run_ngpu.sh N [bench args]
N=${1:-2}; shift
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N "$@" > gpurun_out/bench_${N}gpu_r2.json 2> gpurun_out/bench_${N}gpu_r2.err
echo "rc=$?"; tail -c 900 gpurun_out/bench_${N}gpu_r2.json; tail -3 gpurun_out/bench_${N}gpu_r2.err
