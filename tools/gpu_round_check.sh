# One gpurun call that refreshes what DESIGN.md / profiles/ quote: GPU test suite, default bench, the ncu launch list of a
# bench step (our kernels only) and ncu --set full captures of the two largest kernels.  Usage: gpurun -- 'bash tools/gpu_round_check.sh [quick]'
set -x
KN='regex:fused_preprocess|pack_bits|find_crossings|trace_segments|link_loops|select_quad|homography|cells_from_frames|tc_conv|tc_fc|mask_not_found|reset_kernel'
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests2.log 2>&1; tail -3 gpurun_out/gpu_tests2.log
if [ "$1" = "quick" ]; then
  timeout 600 python bench.py --no-cpu-baseline --no-other-configs > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
else
  timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -c 60 --csv --log-file gpurun_out/launches_all_1024f.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --e2e-frames 64 > gpurun_out/ncu_bench.log 2>&1
if [ "$1" != "quick" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_preprocess_warp -s 2 -c 1 -f -o gpurun_out/k1w python tools/k1_ab.py 256 1 > gpurun_out/ncu_k1.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s 1 -c 1 -f -o gpurun_out/k5conv python tools/k5_ab.py 20736 1 > gpurun_out/ncu_k5.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_fc_tma -s 1 -c 1 -f -o gpurun_out/k5fc python tools/k5_ab.py 82944 1 > gpurun_out/ncu_k5fc.log 2>&1
fi
