#!/bin/bash
# Strong scaling of the headline workload (batch 64 over N GPUs) and weak scaling of the bulk render (1024 voices per
# GPU) on one 8-GPU box; every run under its own timeout.  Writes gpurun_out/r02_scale_n{N}.json / r02_bulk_n{N}.json.
mkdir -p gpurun_out
port=29600
for n in 1 2 4 8; do
  port=$((port+1))
  if [ $n = 1 ]; then
    timeout 200 python bench.py --gpus 1 --steps 100 --warmup 10 --skip-cpu --skip-configs --skip-kernels > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
  else
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 100 --warmup 10 --skip-kernels > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err
  fi
  echo "train n=$n rc=$?"
done
for n in 1 8; do
  port=$((port+1))
  if [ $n = 1 ]; then
    timeout 200 python bench.py --gpus 1 --workload bulk --steps 5 --warmup 2 > gpurun_out/r02_bulk_n1.json 2> gpurun_out/r02_bulk_n1.err
  else
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload bulk --steps 5 --warmup 2 > gpurun_out/r02_bulk_n$n.json 2> gpurun_out/r02_bulk_n$n.err
  fi
  echo "bulk n=$n rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_scale_n*.json")) + sorted(glob.glob("gpurun_out/r02_bulk_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"], 4), f'{d["value"]:.4g}', f'e2e {d["e2e"]["value"]:.4g}')
    except Exception as e:
        print(f, "unreadable", e)
PY
