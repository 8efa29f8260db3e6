#!/bin/bash
# A/B of the interactive frame (BASELINE config 5) over library variants and tuning knobs:
#   tools/ab_config5.sh "variant1 variant2 ..." "ENV1=VAL ENV2=VAL ..."
for v in ${1:-base}; do
  if [ "$v" = base ]; then unset LP_LIB_VARIANT; else export LP_LIB_VARIANT=$v; fi
  echo -n "variant $v: "
  timeout 200 python tools/run_configs.py config5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('frame %.3f ms median %.3f p99, denoise %.3f' % (d['frame_ms_median'], d['frame_ms_p99'], d['denoise_ms_median']))"
done
unset LP_LIB_VARIANT
for b in $2; do
  echo -n "env $b: "
  env $b timeout 200 python tools/run_configs.py config5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('frame %.3f ms median %.3f p99, denoise %.3f' % (d['frame_ms_median'], d['frame_ms_p99'], d['denoise_ms_median']))"
done
