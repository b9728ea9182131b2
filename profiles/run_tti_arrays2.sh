#!/bin/bash
# Second pass on the array-parameter TTI kernel: parity test, prefetch on/off, one full ncu capture.
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tti_array" > $OUT/${TAG}_tti_tests.log 2>&1
tail -5 $OUT/${TAG}_tti_tests.log
B="--steps 2 --warmup 3 --nt 32 --no-e2e --no-cpu --no-extra --no-parity"
for pf in 1 0; do
  B2_TTI_ARR_PREFETCH=$pf timeout 300 python bench.py --workload tti-arrays --grid 512 --space-order 8 $B > $OUT/${TAG}_ttiarr_so8_pf$pf.json 2> $OUT/${TAG}_ttiarr_so8_pf$pf.err
  python - <<PY
import json
try:
    d = json.loads(open('$OUT/${TAG}_ttiarr_so8_pf$pf.json').read().strip().splitlines()[-1])
    print('so=8 prefetch=$pf', round(d['value'], 1), 'GPts/s', round(d['ms_per_step'], 2), 'ms/apply', d['roofline'] and round(d['roofline']['launch_ms'], 3), 'ms', d['roofline'] and round(d['roofline']['frac'], 3))
except Exception as e:
    print('failed', e); print(open('$OUT/${TAG}_ttiarr_so8_pf$pf.err').read()[-1500:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tti_fused -s 4 -c 1 -f -o $OUT/${TAG}_ttiarr_so8 \
    python bench.py --workload tti-arrays --grid 512 --space-order 8 --steps 1 --warmup 1 --nt 8 --no-e2e --no-cpu --no-extra --no-parity > /dev/null 2>&1
ncu -i $OUT/${TAG}_ttiarr_so8.ncu-rep --page details > $OUT/${TAG}_ttiarr_so8_ncu_details.txt 2>/dev/null
ncu -i $OUT/${TAG}_ttiarr_so8.ncu-rep --page raw --csv > $OUT/${TAG}_ttiarr_so8_ncu_raw.csv 2>/dev/null
ls -la $OUT | grep ${TAG}
