#!/bin/bash
# ncu --set full capture of each stage kernel on the bench workload (one lane, one stream), and the launch
# list of one bench step.  Run under gpurun from the repo root: bash profiles/ncu_capture.sh <tag>
# (seam / simplifier kernels: profiles/ncu_capture_widened.sh)
TAG=${1:-r02}
export LVN_LANES=1 LVN_STREAMS=1
for K in k_hermite_terrain k_leaves k_solve k_rows k_columns; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f \
      -o gpurun_out/prof_${K}_${TAG} python profiles/one_batch.py > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
unset LVN_LANES LVN_STREAMS
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --region-seconds 0.001 > gpurun_out/launches_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -12
