#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (both arms), then the ncu passes (launch list + full capture).
# usage: scripts/gpu_round.sh [tag]
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${tag}_pytest_gpu.log
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
echo "reference exit $?"; cat gpurun_out/${tag}_bench_reference.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; cat gpurun_out/${tag}_bench.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --compute fp32 > gpurun_out/${tag}_bench_fp32.json 2> gpurun_out/${tag}_bench_fp32.err
echo "bench fp32 exit $?"; cat gpurun_out/${tag}_bench_fp32.json
BCMD="python bench.py --steps 2 --warmup 1 --no-cpu --updates 32"
$BCMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $BCMD > gpurun_out/${tag}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
$BCMD > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_mcmc -s 1 -c 2 -f -o gpurun_out/${tag}_prof $BCMD > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out | tail -20
