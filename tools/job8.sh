cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job8.log; rm -f $L
echo "=== attention + kernel + block tests" >> $L
timeout 240 python -m pytest tests/test_attention_tc_gpu.py tests/test_kernels_gpu.py tests/test_blocks_gpu.py tests/test_lm_block_gpu.py -q 2>&1 | grep -v "^  \|^E    \|^$\|^The new\|^Flamingo" | cut -c1-300 | tail -15 >> $L
echo "=== bench_attn" >> $L
timeout 150 python tools/bench_attn.py --out gpurun_out/r02_bench_attn_v4.json 2>&1 | tail -9 >> $L
echo "=== C5 perceiver" >> $L
timeout 120 python tools/bench_perceiver.py 2>&1 | tail -1 >> $L
echo "=== bench OF-3B N=1" >> $L
timeout 400 python bench.py --gemm-shapes gpurun_out/r02_gemm_by_shape.json > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -5 gpurun_out/r02_bench_n1.err | cut -c1-400 >> $L
python - <<'PY' >> $L
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('of3b', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['clocks'], d['config']['cuda_graph'])
    print('roofline', round(r['achieved'],1), round(r['frac'],3), round(r['gemm_ms_per_step'],2), r['method'][:60])
    print({k:(round(v['TFLOP/s']),round(v['ms_per_step'],2)) for k,v in r['by_variant'].items()})
    print(d.get('gpu_eager_reference'))
except Exception as e:
    print('parse failed', e)
PY
cat $L | cut -c1-1500
