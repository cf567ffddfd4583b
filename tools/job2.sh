cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job2.log; rm -f $L
echo "=== attention BIG" >> $L
timeout 90 python -m pytest tests/test_attention_tc_gpu.py -q -k "media_attention and (case7 or case8 or case9 or case10)" 2>&1 | grep -v "^  \|^E    \|^$" | cut -c1-300 | tail -25 >> $L
echo "=== full gpu suite (minus parity)" >> $L
timeout 300 python -m pytest tests -q -m gpu --deselect tests/test_fullsize_parity_gpu.py 2>&1 | grep -v "^  \|^E    \|^$" | cut -c1-300 | tail -40 >> $L
echo "=== fullsize parity" >> $L
timeout 400 python -m pytest tests/test_fullsize_parity_gpu.py -q -s 2>&1 | grep -v "^  \|^$" | cut -c1-600 | tail -60 >> $L
echo "=== bench" >> $L
timeout 200 python bench.py > gpurun_out/r02_bench_tcattn.json 2> gpurun_out/r02_bench_tcattn.err
python - <<'PY' >> $L
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_tcattn.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['roofline']['frac'])
except Exception as e:
    print('bench parse failed', e)
PY
grep -E "^===|passed|failed|rror|\[OF|\[ViT|^[0-9]" $L | cut -c1-400 | head -80
