cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job15.log; rm -f $L
echo "=== full gpu suite (2 GPUs visible: includes the NCCL test)" >> $L
timeout 600 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|^E    \|^$\|^The new\|^Flamingo" | cut -c1-300 | tail -8 >> $L
echo "=== smoke" >> $L
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | cut -c1-300 >> $L
echo "=== bench N=2 (final defaults)" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err
python - <<'PY' >> $L
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n2_final.json').read().strip().splitlines()[-1])
    print('n2 final tok/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['clocks'], d['config']['cuda_graph'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r02_bench_n2_final.err').read()[-1500:])
PY
echo "=== bench N=1 (final, full line) on GPU 0" >> $L
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --gemm-shapes gpurun_out/r02_gemm_by_shape_final.json > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err
python - <<'PY' >> $L
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n1_final.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('n1 final tok/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['clocks'], 'gemm', round(r['achieved']), round(r['frac'],3), d['config']['cuda_graph'], 'launches', d['gpu_launches'])
    print('   eager ref', {k:(round(v['value']), round(v['ours_over_this'],2)) for k,v in d['gpu_eager_reference'].items() if isinstance(v, dict) and 'value' in v}, 'cpu', round(d['cpu_baseline']['value'],1))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r02_bench_n1_final.err').read()[-1500:])
PY
cat $L | cut -c1-400
