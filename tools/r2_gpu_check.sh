set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gpu_tests.log 2>&1; tail -25 gpurun_out/r2b_gpu_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-other-configs --stream-seconds 0 --parity-frames 32 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo bench rc=$?; tail -3 gpurun_out/r2b_bench.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r2b_bench.json'))
    print("value",j["value"],"e2e",j["e2e"]["value"]); print(j["stage_ms_per_step"]); print(j["parity"]); print(j["parity_coreml_weights"])
except Exception as e: print("bench json:",e)
PY
for k in k4 k5 k5f k1 k2; do timeout 120 python tools/prof_stage.py $k 1024 5; done
timeout 300 python tools/prof_stage.py k4 256 2 > gpurun_out/plain_k4.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:cells_from_frames -s 1 -c 1 -f -o gpurun_out/r2b_k4 python tools/prof_stage.py k4 256 2 > gpurun_out/ncu_k4.log 2>&1; tail -2 gpurun_out/ncu_k4.log
timeout 300 python tools/prof_stage.py k5 256 2 > gpurun_out/plain_k5.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s 1 -c 1 -f -o gpurun_out/r2b_k5conv python tools/prof_stage.py k5 256 2 > gpurun_out/ncu_k5.log 2>&1; tail -2 gpurun_out/ncu_k5.log
