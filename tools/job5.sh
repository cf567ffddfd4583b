cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job5.log; rm -f $L
for sel in "media_attention and (case0 or case1 or case2 or case3)" "media_attention and (case4 or case5 or case6)" "media_attention and (case7 or case8 or case9 or case10)" "uniform_rows or pure_causal" "dense_attention"; do
  echo "=== attn tests: $sel" >> $L
  timeout 75 python -m pytest tests/test_attention_tc_gpu.py -q -x -k "$sel" 2>&1 | grep -v "^  \|^E    \|^$" | cut -c1-300 | tail -12 >> $L
done
echo "=== bench_attn v2 fwd (default)" >> $L
timeout 150 python tools/bench_attn.py --out gpurun_out/r02_bench_attn_v2.json 2>&1 | tail -9 >> $L
echo "=== bench_attn v1 fwd" >> $L
OFK_ATTN_FWD_V1=1 timeout 150 python tools/bench_attn.py --out gpurun_out/r02_bench_attn_v1.json 2>&1 | tail -9 >> $L
echo "=== parity + ddp single gpu" >> $L
timeout 500 python -m pytest tests/test_fullsize_parity_gpu.py tests/test_ddp_nccl_gpu.py -q -s 2>&1 | grep -v "^  \|^$\|^The new\|^Flamingo" | cut -c1-1200 | tail -30 >> $L
echo "=== ncu LM fwd/bwd" >> $L
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_.*_tc_kernel -c 2 -f -o gpurun_out/r02_attn_lm python tools/bench_attn.py --ncu --shapes 5 2>&1 | tail -3 >> $L
cat $L | cut -c1-1200
