cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job7.log; rm -f $L
echo "=== attention + kernel + lm-block tests" >> $L
timeout 200 python -m pytest tests/test_attention_tc_gpu.py tests/test_kernels_gpu.py tests/test_lm_block_gpu.py tests/test_fullsize_properties_gpu.py -q 2>&1 | grep -v "^  \|^E    \|^$\|^The new\|^Flamingo" | cut -c1-300 | tail -15 >> $L
echo "=== bench_attn" >> $L
timeout 150 python tools/bench_attn.py --out gpurun_out/r02_bench_attn_v3.json 2>&1 | tail -9 >> $L
echo "=== C5 perceiver (tc / legacy attention)" >> $L
timeout 120 python tools/bench_perceiver.py 2>&1 | tail -1 >> $L
OFK_ATTN_LEGACY=1 timeout 120 python tools/bench_perceiver.py 2>&1 | tail -1 >> $L
echo "=== bench OF-3B N=1 (full line)" >> $L
timeout 400 python bench.py --gemm-shapes gpurun_out/r02_gemm_by_shape.json > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 3000 gpurun_out/r02_bench_n1.json >> $L
echo "=== bench OF-9B slice N=1" >> $L
timeout 400 python bench.py --model of9b --batch 8 --t_img 5 --t_txt 512 --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-eager-ref > gpurun_out/r02_bench_of9b_n1.json 2> gpurun_out/r02_bench_of9b_n1.err
python - <<'PY' >> $L
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_of9b_n1.json').read().strip().splitlines()[-1])
    print('of9b', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['clocks'], round(d['roofline']['frac'],3))
except Exception as e:
    print('of9b parse failed', e); print(open('gpurun_out/r02_bench_of9b_n1.err').read()[-1500:])
PY
cat $L | cut -c1-3000
