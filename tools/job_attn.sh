cd $GRAFT_REPO_ROOT
rm -f gpurun_out/r02_attn_tc.log
for sel in "media_attention and (case0 or case1 or case2)" "media_attention and (case3 or case4 or case5 or case6)" "media_attention and (case7 or case8 or case9 or case10)" "uniform_rows" "dense_attention and (case0 or case1 or case2 or case3)" "dense_attention and (case4 or case5 or case6)" "pure_causal"; do
  echo "=== $sel" >> gpurun_out/r02_attn_tc.log
  timeout 75 python -m pytest tests/test_attention_tc_gpu.py -q -x -k "$sel" 2>&1 | grep -v "^  \|^E    \|^$" | cut -c1-300 | tail -25 >> gpurun_out/r02_attn_tc.log
done
grep -E "^===|passed|failed|rror" gpurun_out/r02_attn_tc.log | cut -c1-200 | head -60
