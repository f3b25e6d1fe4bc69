mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_recurrent.py tests/test_gpu_statistical.py tests/test_gpu_edges.py -x -q -m gpu > gpurun_out/s12_tests.log 2>&1; tail -3 gpurun_out/s12_tests.log
python tools/bench_cfg.py --cfg 5 2>&1 | tail -1
python tools/bench_cfg.py --cfg 4 2>&1 | tail -1
