cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job12.log; rm -f $L
echo "=== gemm tests (mix default)" >> $L
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_properties_gpu.py tests/test_blocks_gpu.py -q 2>&1 | tail -3 >> $L
echo "=== ncu attention: xattn C2 (0), perceiver C5 (3), ViT (4)" >> $L
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_.*_kernel -c 12 -f -o gpurun_out/r02_attn_shapes python tools/bench_attn.py --ncu --shapes 0,3,4 2>&1 | tail -2 >> $L
echo "=== ncu gemm FFN shapes" >> $L
timeout 300 ncu --set full --clock-control none -k regex:gemm2_kernel -s 6 -c 6 -f -o gpurun_out/r02_gemm2_ffn python tools/ncu_gemm_target.py 2>&1 | tail -2 >> $L
echo "=== launch list of one eager step" >> $L
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python tools/profile_step_plain.py 2>&1 | tail -2 >> $L
cat $L | cut -c1-300
