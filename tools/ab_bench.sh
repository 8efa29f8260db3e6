#!/bin/bash
# A/B measurements on the GPU box: one bench.py line per library variant / tuning knob.
#   tools/ab_bench.sh TAG "variant1 variant2 ..." "ENV1=VAL ENV2=VAL ..."
# Library variants are built beforehand with `python -m loupiote_b200._build --variant NAME -D...`.
TAG=${1:-ab}
VARIANTS=${2:-base}
BLOCKS=${3:-}
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_ab.jsonl
: > $OUT
for v in $VARIANTS; do
  if [ "$v" = base ]; then unset LP_LIB_VARIANT; else export LP_LIB_VARIANT=$v; fi
  echo "{\"variant\": \"$v\"}" >> $OUT
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --job-spp 0 >> $OUT 2>> gpurun_out/${TAG}_ab.err
done
unset LP_LIB_VARIANT
for b in $BLOCKS; do
  echo "{\"env\": \"$b\"}" >> $OUT
  env $b timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --job-spp 0 >> $OUT 2>> gpurun_out/${TAG}_ab.err
done
python - <<PY
import json
for l in open("$OUT"):
    d = json.loads(l)
    if "value" not in d:
        print(d, end=" ")
        continue
    print("%.0f Mrays/s  %.2f ms/step  kernels %s  ext_frac %.3f path_frac %.3f" % (
        d["value"], d["ms_per_step"], {k: round(v / d["steps"], 2) for k, v in d["kernel_ms"].items()},
        d["roofline"]["frac"], d["roofline_fraction_of_path"]))
PY
