mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_recurrent.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_predict.py -x -q -m gpu > gpurun_out/s15_tests.log 2>&1; tail -3 gpurun_out/s15_tests.log
python tools/bench_cfg.py --cfg 4 2>&1 | tail -1
python - <<'PY'
import json, subprocess, sys
sys.path.insert(0, '.')
import bench, torch
import bayesflux_b200 as bf
print(json.dumps(bench.secondary_cfg4(bf, torch, None, 1, 0, 0))[:600])
PY
