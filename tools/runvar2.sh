run() { timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 2 --warmup 1 2>&1 | grep '"metric"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g ms/step %.1f e2e %.4g'%(d['value'],d['ms_per_step'],d['e2e']['value']))"; }
echo baseline; run 29531
cp mezzo_b200/libmezzo_b200.so /tmp/orig.so; cp tools/_var/lib_ldq.so mezzo_b200/libmezzo_b200.so
echo ldq_cached; run 29532
timeout -s KILL 300 python -m pytest tests/test_sim_multi_gpu.py -x -q 2>&1 | tail -2
cp /tmp/orig.so mezzo_b200/libmezzo_b200.so
