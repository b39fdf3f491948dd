#!/bin/bash
# final round-2 measurements on 8 GPUs: bench under torchrun, bench through the native multi-GPU C ABI, C5 default-mode rows
set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_bench_8gpu.json 2> gpurun_out/r02_final_bench_8gpu.err
timeout 150 python bench.py --gpus 8 --native-dist --steps 5 --warmup 3 --no-cpu-baseline --no-cull > gpurun_out/r02_final_bench_8gpu_native.json 2> gpurun_out/r02_final_bench_8gpu_native.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/c5_sweep.py --brute 100000,1000000 --cull "" > gpurun_out/r02_final_c5_8gpu.jsonl 2> gpurun_out/r02_final_c5_8gpu.err
