#!/bin/bash
# time the variant libraries of tools/build_variants.py: one short bench run each, kernel times per step
for f in acropolis_b200/variants/lib_*.so; do
  tag=$(basename $f .so); tag=${tag#lib_}
  ACROPOLIS_B200_LIB=$PWD/$f python bench.py --steps ${STEPS:-2} --warmup 3 --no-cpu > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "value %.0f"%d["value"], "ms/step %.2f"%d["ms_per_step"], {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, {k:round(v,2) for k,v in d["stage_ms_per_step"].items()})
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/ab_$tag.err").read()[-600:])
PY
done
