mkdir -p gpurun_out
for v in gemm_probe gemm_probe_noearly gemm_probe_no224 gemm_probe_base; do
  bad=0
  for i in 1 2 3 4 5 6 7 8; do
    timeout 120 tools/$v.exe 2 1 > gpurun_out/s4_$v.$i.log 2>&1
    n=$(grep -c MISMATCH gpurun_out/s4_$v.$i.log); bad=$((bad+n))
    grep MISMATCH gpurun_out/s4_$v.$i.log | cut -c1-110
  done
  echo "$v: $bad mismatches in 8 runs"
done
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/s4_tests.log 2>&1; tail -3 gpurun_out/s4_tests.log
BCMD="python bench.py --steps 5 --warmup 2 --no-cpu --no-secondary"
BFX_TC_GENERIC=1 $BCMD 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('generic', d['value'], d['e2e']['value'])"
$BCMD 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fixed', d['value'], d['e2e']['value'])"
