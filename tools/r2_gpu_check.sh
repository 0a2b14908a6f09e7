# One gpurun call that refreshes what DESIGN.md / profiles/ quote: GPU test suite, default bench, the ncu launch list of a
# bench step and ncu --set full captures of K2 and K4.  Usage: gpurun -- 'bash tools/r2_gpu_check.sh'
set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2k_gpu_tests.log 2>&1; tail -4 gpurun_out/r2k_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo bench rc=$?; tail -3 gpurun_out/r2k_bench.err
for k in k1 k2 k4 k5; do timeout 120 python tools/prof_stage.py $k 1024 5; done
KN='regex:fused_preprocess|pack_bits|find_crossings|trace_segments|link_loops|select_quad|homography|cells_from_frames|tc_conv|tc_fc|mask_not_found|reset_kernel|rst_scan|huff_idct|color_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -c 80 --csv --log-file gpurun_out/r2k_launches_bench_step.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --stream-seconds 0 --parity-frames 0 --e2e-frames 64 > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/prof_stage.py k2 256 2 > gpurun_out/plain_k2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"trace_segments|select_quad" -s 2 -c 2 -f -o gpurun_out/r2k_k2 python tools/prof_stage.py k2 256 2 > gpurun_out/ncu_k2.log 2>&1; tail -1 gpurun_out/ncu_k2.log
timeout 300 python tools/prof_stage.py k4 256 2 > gpurun_out/plain_k4.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:cells_from_frames -s 1 -c 1 -f -o gpurun_out/r2k_k4 python tools/prof_stage.py k4 256 2 > gpurun_out/ncu_k4.log 2>&1; tail -1 gpurun_out/ncu_k4.log
