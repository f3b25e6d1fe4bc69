mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_recurrent.py -x -q -m gpu > gpurun_out/s16_tests.log 2>&1; tail -2 gpurun_out/s16_tests.log
python tools/lstm_probe.py 2>&1 | tail -1
python - <<'PY'
import json, subprocess, sys
sys.path.insert(0, '.')
import bench, torch
import bayesflux_b200 as bf
r=bench.secondary_cfg4(bf, torch, None, 1, 0, 0); print(r['ms'], r['grad_evals_per_s'])
r=bench.secondary_cfg4(bf, torch, None, 1, 0, 0, chains=256); print('256 chains', r['ms'], r['grad_evals_per_s'])
PY
