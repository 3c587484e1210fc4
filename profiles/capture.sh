#!/bin/bash
# Profile capture of one round (run on the GPU box through gpurun, one GPU):
#   bash profiles/capture.sh <tag>
# writes gpurun_out/<tag>_launches.csv (ncu launch list of `bench.py --steps 2 --warmup 3`, gpu__time_duration only)
# and gpurun_out/<tag>_prof.ncu-rep (`--set full` of the composite + per-Gaussian kernels of one c2 view).
# Numbers printed under ncu are never bench values.
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_composite|k_preprocess' -c 4 \
    -o gpurun_out/${tag}_prof -f python profiles/one_view.py c2 1 > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/${tag}_*
