mkdir -p gpurun_out
for i in 1 2 3; do timeout 120 tools/gemm_probe.exe 2 1 > gpurun_out/s13_probe.$i.log 2>&1; echo "probe rc $? mism $(grep -c MISMATCH gpurun_out/s13_probe.$i.log)"; done
grep -A2 "^cfg3" gpurun_out/s13_probe.1.log | grep -v "^--"
timeout 600 python -m pytest tests/test_gpu_stream.py -x -q -m gpu > gpurun_out/s13_stream.log 2>&1; tail -2 gpurun_out/s13_stream.log
timeout 120 python tools/cfg3_probe.py --steps 100 2>&1 | tail -1
