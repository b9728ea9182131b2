#!/bin/bash
# End-of-round check on ONE GPU: the whole GPU suite, the driver-style bench lines, and the evidence for the one
# kernel that changed since the r2z captures (array-parameter TTI).  Usage: gpurun -- 'bash profiles/run_final.sh r2y'
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
python - <<PY
import json
d = json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d.get('parity_check'))
for k, v in (d.get('configs') or {}).items():
    print(k, round(v['value'], 1), v['roofline'] and round(v['roofline']['frac'], 3))
print(open('$OUT/${TAG}_bench_reference.json').read()[:300])
PY
A="--workload tti-arrays --grid 512 --space-order 8 --steps 1 --warmup 1 --nt 8 --no-e2e --no-cpu --no-extra --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_ttiarr_so8.csv python bench.py $A > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tti_fused -s 4 -c 1 -f -o $OUT/${TAG}_ttiarr_so8 python bench.py $A > /dev/null 2>&1
ncu -i $OUT/${TAG}_ttiarr_so8.ncu-rep --page details > $OUT/${TAG}_ttiarr_so8_ncu_details.txt 2>/dev/null
ncu -i $OUT/${TAG}_ttiarr_so8.ncu-rep --page raw --csv > $OUT/${TAG}_ttiarr_so8_ncu_raw.csv 2>/dev/null
ls -la $OUT | grep ${TAG}
