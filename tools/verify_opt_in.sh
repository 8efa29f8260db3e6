#!/bin/bash
# The opt-in features that were validated on the host only: run their
# gated GPU tests and their A/B lines on ONE B200.  About 40 s of GPU time.
#
#   gpurun --timeout 150 -- 'bash tools/verify_opt_in.sh'
#
# Outputs: gpurun_out/opt_in_tests.log, gpurun_out/opt_in_bench.jsonl
set -u
mkdir -p gpurun_out
LP_TEST_LBVH_TREELETS=1 LP_TEST_LBVH_BLOCK_TLAS=1 timeout 90 python -m pytest \
  tests/test_gpu_lbvh.py tests/test_gpu_zz_cli_device_build.py -q --tb=short -p no:cacheprovider \
  > gpurun_out/opt_in_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/opt_in_tests.log
rm -f gpurun_out/opt_in_bench.jsonl
for args in "--treelets 0" "--treelets 1" "--treelets 2" "--treelets 2 --block-tlas"; do
  timeout 40 python tools/lbvh_bench.py --spp 32 --steps 2 --build-repeats 5 $args \
    --out gpurun_out/opt_in_bench.jsonl > /dev/null 2>> gpurun_out/opt_in_tests.log
done
tail -3 gpurun_out/opt_in_tests.log
grep '"builder": "lbvh"' gpurun_out/opt_in_bench.jsonl | cut -c1-400
