#!/bin/bash
# ncu --set full capture of the seam and simplifier kernels on their bench workloads: bash profiles/ncu_capture_widened.sh <tag>
TAG=${1:-r02}
ncu --set full --clock-control none --import-source on -k regex:k_seam$ -s 2 -c 1 -f \
    -o gpurun_out/prof_k_seam_${TAG} python profiles/bench_seams.py > gpurun_out/ncu_k_seam_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_simplify -s 2 -c 1 -f \
    -o gpurun_out/prof_k_simplify_${TAG} python profiles/bench_simplify.py ring > gpurun_out/ncu_k_simplify_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_simplify|k_pack|k_seam" -c 40 --csv \
    --log-file gpurun_out/launches_update_${TAG}.csv python profiles/bench_clipmap_update.py > /dev/null 2>&1
