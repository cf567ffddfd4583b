cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job4.log; rm -f $L
echo "=== bench_attn (v1 kernels)" >> $L
timeout 200 python tools/bench_attn.py --out gpurun_out/r02_bench_attn_v1.json 2>&1 | tail -12 >> $L
echo "=== parity" >> $L
timeout 500 python -m pytest tests/test_fullsize_parity_gpu.py -q -s 2>&1 | grep -v "^  \|^$\|^The new\|^Flamingo" | cut -c1-900 | tail -30 >> $L
echo "=== ncu LM fwd/bwd" >> $L
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_.*_tc_kernel -c 2 -f -o gpurun_out/r02_attn_lm_v1 python tools/bench_attn.py --ncu --shapes 5 2>&1 | tail -3 >> $L
cat $L | cut -c1-900
