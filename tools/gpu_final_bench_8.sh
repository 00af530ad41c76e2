#!/bin/bash
# the 8-GPU lines kept under profiles/: weak scaling of the default workload, the whole scan as one call, the 4x-resolution scan
set -x
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$R --master-port 29511 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/bench_8gpu.err
$R --master-port 29512 bench.py --gpus 8 --config cfg3scan > gpurun_out/r02_bench_cfg3scan_8gpu.json 2> gpurun_out/bench_cfg3scan_8gpu.err
[ -n "$SKIP_CFG5" ] || $R --master-port 29513 bench.py --gpus 8 --config cfg5 --steps 2 > gpurun_out/r02_bench_cfg5_8gpu.json 2> gpurun_out/bench_cfg5_8gpu.err
tail -c 300 gpurun_out/r02_bench_8gpu.json
