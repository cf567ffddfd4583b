cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job9.log; rm -f $L
run_bench() {  # tag, ngpu, extra args...
  tag=$1; n=$2; shift 2
  echo "=== bench $tag (N=$n) $*" >> $L
  timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --warmup 3 --no-cpu-baseline --no-gpu-eager-ref "$@" > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err
  python - "$tag" <<'PY' >> $L
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02_bench_{n}.json').read().strip().splitlines()[-1])
    print(n, 'tok/s', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['clocks'], 'gemm frac', round(d['roofline']['frac'],3), 'graph', d['config']['cuda_graph'])
except Exception as e:
    print(n, 'bench parse failed', e); print(open(f'gpurun_out/r02_bench_{n}.err').read()[-1200:])
PY
}
run_bench n8 8 --steps 8
CUDA_VISIBLE_DEVICES=0,1,2,3 run_bench n4 4 --steps 8
run_bench of9b_n8 8 --model of9b --batch 8 --t_img 5 --t_txt 512 --steps 4
nvidia-smi topo -m 2>/dev/null | head -12 >> $L
cat $L | cut -c1-500
