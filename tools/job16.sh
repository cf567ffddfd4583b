cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job16.log; rm -f $L
for r in 16 0; do
echo "=== bench N=8 reserve=$r" >> $L
OFK_COMM_RESERVE_SMS=$r timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2957$((r/16)) bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-ref > gpurun_out/r02_bench_n8_final_r$r.json 2> gpurun_out/r02_bench_n8_final_r$r.err
python - $r <<'PY' >> $L
import json,sys
r=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02_bench_n8_final_r{r}.json').read().strip().splitlines()[-1])
    print('n8 reserve', r, 'tok/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['clocks'], d['config']['cuda_graph'])
except Exception as e:
    print('parse failed', e); print(open(f'gpurun_out/r02_bench_n8_final_r{r}.err').read()[-1500:])
PY
done
cat $L | cut -c1-300
