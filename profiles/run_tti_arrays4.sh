#!/bin/bash
# Fourth pass: the three-plane factor ring (B2_TTI_ARR_CT=2: stage B reads cy / cz / cx from shared memory too).
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
B2_TTI_ARR_CT=2 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tti_array" > $OUT/${TAG}_tti_tests_ct2.log 2>&1
tail -5 $OUT/${TAG}_tti_tests_ct2.log
B="--steps 2 --warmup 3 --nt 32 --no-e2e --no-cpu --no-extra --no-parity"
for cfg in "8 1" "8 2" "4 1" "4 2"; do
  set -- $cfg
  B2_TTI_ARR_CT=$2 timeout 300 python bench.py --workload tti-arrays --grid 512 --space-order $1 $B > $OUT/${TAG}_ttiarr_so$1_ct$2.json 2> $OUT/${TAG}_ttiarr_so$1_ct$2.err
  python - <<PY
import json
try:
    d = json.loads(open('$OUT/${TAG}_ttiarr_so$1_ct$2.json').read().strip().splitlines()[-1])
    print('so=$1 ct=$2', round(d['value'], 1), 'GPts/s', d['roofline'] and round(d['roofline']['launch_ms'], 3), 'ms', d['roofline'] and round(d['roofline']['frac'], 3))
except Exception as e:
    print('so=$1 ct=$2 failed', e); print(open('$OUT/${TAG}_ttiarr_so$1_ct$2.err').read()[-1500:])
PY
done
