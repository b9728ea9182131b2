#!/bin/bash
# Third pass on the array-parameter TTI kernel: the TMA-staged factor tiles (B2_TTI_ARR_CT) and the L1 no-allocate hint
# (B2_TTI_ARR_HINT) — parity first, then the timing matrix.
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
B2_TTI_ARR_CT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tti_array" > $OUT/${TAG}_tti_tests_ct.log 2>&1
tail -5 $OUT/${TAG}_tti_tests_ct.log
B2_TTI_ARR_HINT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tti_array" > $OUT/${TAG}_tti_tests_hint.log 2>&1
tail -2 $OUT/${TAG}_tti_tests_hint.log
B="--steps 2 --warmup 3 --nt 32 --no-e2e --no-cpu --no-extra --no-parity"
for cfg in "8 0 0" "8 0 1" "8 1 0" "8 1 1" "4 0 0" "4 1 0" "4 1 1"; do
  set -- $cfg
  B2_TTI_ARR_CT=$2 B2_TTI_ARR_HINT=$3 timeout 300 python bench.py --workload tti-arrays --grid 512 --space-order $1 $B > $OUT/${TAG}_ttiarr_so$1_ct$2_h$3.json 2> $OUT/${TAG}_ttiarr_so$1_ct$2_h$3.err
  python - <<PY
import json
try:
    d = json.loads(open('$OUT/${TAG}_ttiarr_so$1_ct$2_h$3.json').read().strip().splitlines()[-1])
    print('so=$1 ct=$2 hint=$3', round(d['value'], 1), 'GPts/s', round(d['ms_per_step'], 2), 'ms/apply', d['roofline'] and round(d['roofline']['launch_ms'], 3), 'ms', d['roofline'] and round(d['roofline']['frac'], 3))
except Exception as e:
    print('so=$1 ct=$2 hint=$3 failed', e); print(open('$OUT/${TAG}_ttiarr_so$1_ct$2_h$3.err').read()[-1500:])
PY
done
