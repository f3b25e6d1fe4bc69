mkdir -p gpurun_out
for z in 0 1; do
  if [ $z = 1 ]; then export BFX_NO_ZEROCOPY=1; else unset BFX_NO_ZEROCOPY; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$z bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nozc=$z', d['value'], d['e2e']['value'])"
done
