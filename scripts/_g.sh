mkdir -p gpurun_out
T=r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref exit $?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary --compute fp32 > gpurun_out/${T}_bench_fp32.json 2> /dev/null; echo "fp32 exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench.json'))
print('cfg2', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
for k in ('cfg3','cfg5','cfg4','cfg1'):
    v=d['secondary'][k]; print(k, v.get('ms_per_update') or v.get('full_batch_grad_evals_per_s') or v.get('grad_evals_per_s'))
print('fp32', json.load(open('gpurun_out/r02_bench_fp32.json'))['value'])
PY
for c in tc fp32; do python tools/bench_samplers.py --compute $c > gpurun_out/${T}_samplers_$c.json 2>/dev/null; tail -c 400 gpurun_out/${T}_samplers_$c.json; echo; done
LIBBFX=tools/libbfx_timing.so python tools/tc_timing.py > gpurun_out/${T}_phase.txt 2>&1; head -2 gpurun_out/${T}_phase.txt
BCMD="python bench.py --steps 2 --warmup 1 --no-cpu --updates 32 --no-secondary"
$BCMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $BCMD > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_mcmc -s 1 -c 1 -f -o gpurun_out/${T}_prof $BCMD > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full exit $?"
