#!/bin/bash
# bench.py at N GPUs (both metrics), like the driver's scaling run; writes gpurun_out/scale_*.json
N=$1
O=gpurun_out; mkdir -p $O
if [ "$N" = "1" ]; then
  timeout -s KILL 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $O/scale_sim_n1.json 2> $O/scale_sim_n1.err
  timeout -s KILL 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --workload sssp_grid100 > $O/scale_sssp_n1.json 2> $O/scale_sssp_n1.err
else
  timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 3 --warmup 3 > $O/scale_sim_n$N.json 2> $O/scale_sim_n$N.err
  timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 3 --warmup 3 --workload sssp_grid100 > $O/scale_sssp_n$N.json 2> $O/scale_sssp_n$N.err
fi
for f in sim sssp; do tail -1 $O/scale_${f}_n$N.json | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$f N=$N value %.4g ms/step %.1f e2e %.4g'%(d['value'],d['ms_per_step'],d['e2e']['value']))
except Exception as e: print('$f N=$N FAILED', e)"; done
tail -3 $O/scale_sim_n$N.err
