cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job6.log; rm -f $L
echo "=== full gpu suite (2 GPUs visible)" >> $L
timeout 420 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|^E    \|^$\|^The new\|^Flamingo" | cut -c1-400 | tail -30 >> $L
run_bench() {  # name, reserve
  echo "=== bench N=2 $1" >> $L
  env OFK_COMM_RESERVE_SMS=$2 ${3:+NCCL_MAX_CTAS=$3} timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager-ref > gpurun_out/r02_bench_n2_$1.json 2> gpurun_out/r02_bench_n2_$1.err
  python - "$1" <<'PY' >> $L
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02_bench_n2_{n}.json').read().strip().splitlines()[-1])
    print(n, round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['clocks'], round(d['roofline']['frac'],3), d['config']['cuda_graph'])
except Exception as e:
    print(n, 'bench parse failed', e); print(open(f'gpurun_out/r02_bench_n2_{n}.err').read()[-1500:])
PY
}
run_bench reserve0 0
run_bench reserve8 8 8
run_bench reserve16 16 16
cat $L | cut -c1-600
