mkdir -p gpurun_out
python tools/lstm_value_probe.py 148 > gpurun_out/r02_lstm_value_probe.log 2>&1; tail -1 gpurun_out/r02_lstm_value_probe.log
ncu --set full --clock-control none -k regex:k_eval -s 1 -c 1 -f -o gpurun_out/r02_lstm_value python tools/lstm_value_probe.py 148 > gpurun_out/r02_lstm_value_ncu.log 2>&1; echo "ncu value exit $?"
ncu --set full --clock-control none -k regex:k_mcmc -s 1 -c 1 -f -o gpurun_out/r02_lstm_grad python tools/lstm_probe.py 148 4 > gpurun_out/r02_lstm_grad_ncu.log 2>&1; echo "ncu grad exit $?"
