# One gpurun call that refreshes what DESIGN.md / profiles/ quote: GPU test suite, default bench, the ncu launch list of a
# bench step (our kernels only) and ncu --set full captures of the largest kernels.  Usage: gpurun -- 'bash tools/gpu_round_check.sh [quick] [tag]'
set -x
TAG=${2:-r2s}
KN='regex:fused_preprocess|pack_bits|find_crossings|trace_segments|link_loops|select_quad|homography|cells_from_frames|tc_conv|tc_fc|mask_not_found|reset_kernel'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -3 gpurun_out/${TAG}_gpu_tests.log
if [ "$1" = "quick" ]; then
  timeout 600 python bench.py --no-cpu-baseline --no-other-configs > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
else
  timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -c 60 --csv --log-file gpurun_out/${TAG}_launches_bench_step.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --e2e-frames 64 --stream-seconds 0 --parity-frames 0 > gpurun_out/ncu_bench.log 2>&1
if [ "$1" != "quick" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_preprocess_warp -s 2 -c 1 -f -o gpurun_out/${TAG}_k1w python tools/k1_ab.py 256 1 > gpurun_out/ncu_k1.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:cells_from_frames -s 1 -c 1 -f -o gpurun_out/${TAG}_k4 python tools/prof_stage.py k4 256 2 > gpurun_out/ncu_k4.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s 1 -c 1 -f -o gpurun_out/${TAG}_k5conv python tools/k5_ab.py 20736 1 > gpurun_out/ncu_k5.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${TAG}_v3_launches.csv python tools/prof_v3.py 2368 1 > gpurun_out/ncu_v3.log 2>&1
fi
