#!/bin/bash
# Round-2 evidence capture (run on the GPU box through gpurun, one GPU):  bash profiles/capture_r02.sh
# Everything lands in gpurun_out/r02_*; the summaries are then copied into profiles/ by hand.
# Numbers printed under ncu / compute-sanitizer are never bench values.
mkdir -p gpurun_out
# 1. launch list of the bench command (per-kernel share of the step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r02_launches_bench.log 2>&1
# 2. full capture of the composite + per-Gaussian kernels of one c2 view
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_composite|k_preprocess' -c 4 \
    -o gpurun_out/r02_prof -f python profiles/one_view.py c2 1 > gpurun_out/r02_ncu.log 2>&1
# 3. mask kernel (tcgen05) full capture
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_mask_apply' -c 1 \
    -o gpurun_out/r02_mask -f python profiles/mask_one.py 16 > gpurun_out/r02_mask_ncu.log 2>&1
# 4. work counters (instrumented twin of the library)
GOI_RASTER_LIB=goi-hyperplane_b200/lib/libgoi_raster_stats.so timeout 300 python profiles/work_counters.py c2 \
    > gpurun_out/r02_work_counters_c2.json 2> gpurun_out/r02_work_counters.err
# 5. compute-sanitizer on a config-1-sized forward + backward + mask
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python profiles/sanitize_small.py > gpurun_out/r02_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python profiles/sanitize_small.py > gpurun_out/r02_racecheck.log 2>&1
# 6. the other BASELINE configs on one GPU + mask / loss benches
for c in c3 c5_4 c5_8 c5_16 c5_32 c5_64; do
    timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_$c.json
done
timeout 300 python bench.py --impl reference --config c3 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02_c3_reference.json
timeout 300 python bench.py --impl reference --config c5_32 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02_c5_32_reference.json
timeout 300 python profiles/mask_bench.py c2 2>/dev/null | tail -1 > gpurun_out/r02_mask_bench_c2.json
timeout 300 python profiles/mask_bench.py c3 2>/dev/null | tail -1 > gpurun_out/r02_mask_bench_c3.json
ls -la gpurun_out/r02_*
