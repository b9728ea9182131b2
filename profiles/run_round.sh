#!/bin/bash
# One gpurun call that collects everything a round needs on ONE GPU (usage, from the repo root):
#
#   gpurun --timeout 1700 -- 'bash profiles/run_round.sh r2z'
#
# Everything lands in gpurun_out/<tag>_*; copy what is worth judging into profiles/ afterwards.
# Numbers printed by runs under ncu are never bench values (B200_PROFILING.md).
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
set -x

# 1. parity first (the whole GPU suite, as the driver runs it)
python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log

# 2. the bench line (not under a profiler): headline + parity_check + C3/C4/C5 side runs + e2e + CPU arm sample
python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
cut -c1-300 $OUT/${TAG}_bench.json

# 3. launch lists of the same command (share of the step per kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_iso_so8.csv \
    python bench.py --steps 1 --warmup 1 --nt 8 --no-e2e --no-cpu --no-extra --no-parity > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_tti_so8.csv \
    python bench.py --workload tti --grid 768 --steps 1 --warmup 1 --nt 8 --no-e2e --no-cpu --no-extra --no-parity > /dev/null 2>&1

# 4. one full capture per hot kernel (skip the first launches: cold caches / one-off set-up)
ncu --set full --clock-control none --import-source on -k regex:k_iso_tma -s 4 -c 1 -f -o $OUT/${TAG}_iso_so8 \
    python bench.py --steps 1 --warmup 1 --nt 8 --no-e2e --no-cpu --no-extra --no-parity > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tti_ws -s 4 -c 1 -f -o $OUT/${TAG}_tti_so8 \
    python bench.py --workload tti --grid 768 --steps 1 --warmup 1 --nt 8 --no-e2e --no-cpu --no-extra --no-parity > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_iso_tma -s 4 -c 1 -f -o $OUT/${TAG}_iso_so12 \
    python bench.py --space-order 12 --steps 1 --warmup 1 --nt 8 --no-e2e --no-cpu --no-extra --no-parity > /dev/null 2>&1
for k in iso_so8 tti_so8 iso_so12; do
    ncu -i $OUT/${TAG}_$k.ncu-rep --page details > $OUT/${TAG}_${k}_ncu_details.txt 2>/dev/null
    ncu -i $OUT/${TAG}_$k.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_ncu_raw.csv 2>/dev/null
done
ls -la $OUT | tail -20
