cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job13.log; rm -f $L
echo "=== DDP timeline N=2, reserve 16 SMs + NCCL_MAX_CTAS=16" >> $L
OFK_COMM_RESERVE_SMS=16 NCCL_MAX_CTAS=16 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/profile_ddp.py 2>&1 | grep -v "^The new\|^Flamingo\|Warning\|warn\|^\*\*\*\|OMP_NUM" | tail -28 >> $L
echo "=== DDP timeline N=2, default" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/profile_ddp.py 2>&1 | grep -v "^The new\|^Flamingo\|Warning\|warn\|^\*\*\*\|OMP_NUM" | tail -28 >> $L
echo "=== bench N=2 --micro-batches 2 (the reference's LAION + MMC4 step shape)" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 6 --warmup 3 --micro-batches 2 --no-cpu-baseline --no-gpu-eager-ref > gpurun_out/r02_bench_n2_mb2.json 2> gpurun_out/r02_bench_n2_mb2.err
python - <<'PY' >> $L
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n2_mb2.json').read().strip().splitlines()[-1])
    print('n2 mb2 tok/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['clocks'], d['config']['micro_batches'], d['config']['cuda_graph'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r02_bench_n2_mb2.err').read()[-1500:])
PY
cat $L | cut -c1-300
