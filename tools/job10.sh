cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job10.log; rm -f $L
export CUDA_VISIBLE_DEVICES=0
echo "=== gemm + full suite (12 epilogue warps)" >> $L
timeout 400 python -m pytest tests -q -m gpu --deselect tests/test_fullsize_parity_gpu.py 2>&1 | grep -v "^  \|^E    \|^$\|^The new\|^Flamingo" | cut -c1-300 | tail -12 >> $L
for v in ew8 "" ew8 ""; do
  echo "=== bench N=1 variant='$v'" >> $L
  OFK_LIB_VARIANT=$v timeout 300 python bench.py --no-cpu-baseline --no-gpu-eager-ref --steps 8 --gemm-shapes gpurun_out/r02_gemm_by_shape_${v:-ew12}.json > gpurun_out/r02_bench_ab_${v:-ew12}.json 2> gpurun_out/r02_bench_ab_${v:-ew12}.err
  python - "${v:-ew12}" <<'PY' >> $L
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02_bench_ab_{n}.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print(n, 'tok/s', round(d['value']), 'ms', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'gemm', round(r['achieved']), round(r['frac'],3), round(r['gemm_ms_per_step'],2))
    print('   ', {k:(round(v['TFLOP/s']),round(v['ms_per_step'],2)) for k,v in r['by_variant'].items()})
except Exception as e:
    print(n, 'parse failed', e); print(open(f'gpurun_out/r02_bench_ab_{n}.err').read()[-800:])
PY
done
unset CUDA_VISIBLE_DEVICES
echo "=== DDP timeline N=2" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/profile_ddp.py 2>&1 | grep -v "^The new\|^Flamingo\|Warning\|warn" | tail -45 >> $L
cat $L | cut -c1-400
