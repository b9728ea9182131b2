#!/bin/bash
# Array-parameter TTI (k_tti_fused<.., ARR>) on one GPU: parity tests, then the kernel against the two-pass generic
# path, then the plugin's system executor on the reference's elastic example.  Usage: gpurun -- 'bash profiles/run_tti_arrays.sh <tag>'
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tti" > $OUT/${TAG}_tti_tests.log 2>&1
tail -15 $OUT/${TAG}_tti_tests.log
B="--steps 2 --warmup 3 --nt 32 --no-e2e --no-cpu --no-extra --no-parity"
for so in 8 4; do
  timeout 300 python bench.py --workload tti-arrays --grid 512 --space-order $so $B > $OUT/${TAG}_ttiarr_so${so}_fused.json 2> $OUT/${TAG}_ttiarr_so${so}_fused.err
  B2_TTI_ARR_FUSED=0 timeout 300 python bench.py --workload tti-arrays --grid 512 --space-order $so $B > $OUT/${TAG}_ttiarr_so${so}_twopass.json 2> $OUT/${TAG}_ttiarr_so${so}_twopass.err
  for v in fused twopass; do python - <<PY
import json
try:
    d = json.loads(open('$OUT/${TAG}_ttiarr_so${so}_$v.json').read().strip().splitlines()[-1])
    print('so=$so $v', round(d['value'], 1), 'GPts/s', d['roofline'] and round(d['roofline']['launch_ms'], 3), 'ms', d['roofline'] and round(d['roofline']['frac'], 3))
except Exception as e:
    print('so=$so $v failed', e); print(open('$OUT/${TAG}_ttiarr_so${so}_$v.err').read()[-1500:])
PY
  done
done
ROOT=$(pwd)
PYTHONPATH=$ROOT:$ROOT/oracle/refshim:$ROOT/baseline/_ref DEVITO_ARCH=gcc DEVITO_LOGGING=ERROR timeout 600 python profiles/micro/elastic_perf.py 200 > $OUT/${TAG}_elastic_perf.txt 2>&1
tail -4 $OUT/${TAG}_elastic_perf.txt
