#!/bin/bash
# the bench lines kept under profiles/ (one GPU): default arm, reference arm, the other BASELINE configurations
set -x
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/bench_reference.err
for c in cfg1 cfg2 cfg3b cfg4 cfg3scan; do
  python bench.py --config $c --no-cpu > gpurun_out/r02_bench_$c.json 2> gpurun_out/bench_$c.err
done
python bench.py --config cfg5 --no-cpu --steps 2 > gpurun_out/r02_bench_cfg5.json 2> gpurun_out/bench_cfg5.err
tail -c 600 gpurun_out/r02_bench_default.json
