mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_statistical.py -x -q -m gpu > gpurun_out/s11_tests.log 2>&1; tail -3 gpurun_out/s11_tests.log
python tools/bench_cfg.py --cfg 5 2>&1 | tail -2
python bench.py --steps 8 --warmup 3 --no-cpu --no-secondary 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2', d['value'], d['e2e']['value'])"
LIBBFX=tools/libbfx_timing.so python tools/tc_timing.py --sampler hmc --chains 592 --updates 5 > gpurun_out/s11_phase_hmc.txt 2>&1; cat gpurun_out/s11_phase_hmc.txt | head -12
