timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -1
for i in 1 2; do python bench.py --steps 8 --warmup 3 --no-cpu --no-secondary 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2', d['value'], d['e2e']['value'])"; done
