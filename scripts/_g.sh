mkdir -p gpurun_out
T=r02
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref exit $?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${T}_bench.err
python tools/lstm_probe.py > gpurun_out/${T}_lstm_probe.log 2>&1; tail -2 gpurun_out/${T}_lstm_probe.log
CCMD="python tools/cfg3_probe.py --steps 20"
$CCMD > gpurun_out/${T}_cfg3_plain.log 2>&1; tail -1 gpurun_out/${T}_cfg3_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${T}_cfg3_launches.csv $CCMD > gpurun_out/${T}_cfg3_ncu_launches.log 2>&1; echo "ncu cfg3 launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm2 -s 30 -c 8 -f -o gpurun_out/${T}_gemm $CCMD > gpurun_out/${T}_cfg3_ncu_full.log 2>&1; echo "ncu gemm exit $?"
