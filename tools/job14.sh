cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job14.log; rm -f $L
echo "=== train-step / ddp / checkpoint / blocks tests" >> $L
timeout 300 python -m pytest tests/test_train_step_gpu.py tests/test_ddp_nccl_gpu.py tests/test_blocks_gpu.py -q 2>&1 | grep -v "^  \|^E    \|^$\|^The new\|^Flamingo" | cut -c1-300 | tail -10 >> $L
echo "=== DDP timeline N=2 (defaults: 16 SMs reserved, resampler chunked per layer)" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/profile_ddp.py 2>&1 | grep -v "^The new\|^Flamingo\|Warning\|warn\|^\*\*\*\|OMP_NUM" | tail -34 >> $L
for r in 16 0; do
echo "=== bench N=2 reserve=$r" >> $L
OFK_COMM_RESERVE_SMS=$r timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-ref > gpurun_out/r02_bench_n2_final_r$r.json 2> gpurun_out/r02_bench_n2_final_r$r.err
python - $r <<'PY' >> $L
import json,sys
r=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02_bench_n2_final_r{r}.json').read().strip().splitlines()[-1])
    print('n2 reserve', r, 'tok/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['clocks'], d['config']['cuda_graph'])
except Exception as e:
    print('parse failed', e); print(open(f'gpurun_out/r02_bench_n2_final_r{r}.err').read()[-1500:])
PY
done
cat $L | cut -c1-300
