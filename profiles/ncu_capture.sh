#!/bin/bash
# ncu --set full capture of each stage kernel on the bench workload (one lane, one stream), of the
# seam and simplifier kernels on their bench workloads, and the launch list of one bench step.
# Run under gpurun from the repo root: bash profiles/ncu_capture.sh <tag>
TAG=${1:-r01}
export LVN_LANES=1 LVN_STREAMS=1
for K in k_hermite_terrain k_leaves k_rows k_columns; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f \
      -o gpurun_out/prof_${K}_${TAG} python profiles/one_batch.py > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
unset LVN_LANES LVN_STREAMS
ncu --set full --clock-control none --import-source on -k regex:k_seam$ -s 2 -c 1 -f \
    -o gpurun_out/prof_k_seam_${TAG} python profiles/bench_seams.py > gpurun_out/ncu_k_seam_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_simplify -s 2 -c 1 -f \
    -o gpurun_out/prof_k_simplify_${TAG} python profiles/bench_simplify.py ring > gpurun_out/ncu_k_simplify_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_simplify|k_pack|k_seam" -c 40 --csv \
    --log-file gpurun_out/launches_update_${TAG}.csv python profiles/bench_clipmap_update.py > /dev/null 2>&1
ls -la gpurun_out/ | tail -20
