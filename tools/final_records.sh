#!/bin/bash
# The records of a kernel version, on a 1-GPU box:  tools/final_records.sh TAG
#   GPU suite, compute-sanitizer (memcheck + racecheck) over the parity tests, the ncu launch
#   list and physical counters of the bench command, then the bench line itself.
TAG=${1:-vX}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 > $O/${TAG}_gpu_tests.log
cat $O/${TAG}_gpu_tests.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "first_hit or path_trace or empty or multi_wave" 2>&1 | tail -4 > $O/${TAG}_sanitizer_memcheck_parity.log
cat $O/${TAG}_sanitizer_memcheck_parity.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "first_hit_spheres_small or path_trace_cornell" 2>&1 | tail -4 > $O/${TAG}_sanitizer_racecheck_parity.log
cat $O/${TAG}_sanitizer_racecheck_parity.log
METRICS=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_physical; print(ncu_physical.METRICS)")
timeout 900 ncu --metrics $METRICS --clock-control none --kernel-name-base demangled -k regex:"extend4_kernel|trace_pool_kernel|shade_kernel" -s 24 -c 24 --csv --log-file $O/${TAG}_physical.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/${TAG}_physical_bench.json 2> $O/${TAG}_physical.err
python tools/ncu_physical.py $O/${TAG}_physical.csv $O/${TAG}_physical_bench.json profiles/${TAG}_physical_counters.csv && cp profiles/ncu_traffic.json $O/${TAG}_ncu_traffic.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 26 -c 27 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.log
python -c "
import json; d=json.load(open('$O/${TAG}_bench.json'))
for k in ('value','ms_per_step','e2e','parity_check','roofline_fraction_of_path','config5','job','cpu_baseline'): print(k, d.get(k))
print({k: d['roofline'][k] for k in ('achieved','frac','traffic','launch_ms')}); print(json.dumps(d['roofline'].get('physical'))[:600])"
