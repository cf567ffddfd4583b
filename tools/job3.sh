cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job3.log; rm -f $L
echo "=== exp_gate" >> $L
timeout 300 python tools/exp_gate.py 2>&1 | grep -v "^The new\|^Flamingo" | tail -40 >> $L
echo "=== profile tc" >> $L
timeout 200 python tools/profile_step.py 2>&1 | grep -v "^The new\|^Flamingo" | head -40 >> $L
echo "=== profile legacy attention" >> $L
OFK_ATTN_LEGACY=1 timeout 200 python tools/profile_step.py 2>&1 | grep -v "^The new\|^Flamingo" | head -34 >> $L
echo "=== bench A/B (legacy, tc)" >> $L
OFK_ATTN_LEGACY=1 timeout 200 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r02_bench_ab_legacy.json 2>/dev/null
timeout 200 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r02_bench_ab_tc.json 2>/dev/null
python - <<'PY' >> $L
import json
for n in ('legacy','tc'):
    try:
        d=json.loads(open(f'gpurun_out/r02_bench_ab_{n}.json').read().strip().splitlines()[-1])
        print(n, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'])
    except Exception as e:
        print(n, 'bench parse failed', e)
PY
cat $L | cut -c1-250
