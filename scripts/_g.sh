mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n8.json'))
print(d['value'], d['e2e']['value'])
c=d['secondary']['cfg3']
print({k:c.get(k) for k in ('ms_per_update','ms_per_update_bucketed','ms_per_update_nccl_allreduce','peer_ranks','replicas_bit_identical','finite','observations_per_s','error')})
PY
