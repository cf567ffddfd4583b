cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_job11.log; rm -f $L
echo "=== gemm tests (default build)" >> $L
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_properties_gpu.py -q -k "gemm" 2>&1 | tail -3 >> $L
for v in ew8 "" mix ew16 ew8 ""; do
  echo "=== gemm step variant='${v:-default12}'" >> $L
  OFK_LIB_VARIANT=$v timeout 120 python tools/bench_gemm_step.py profiles/r02_gemm_by_shape.json 2>&1 | tail -1 >> $L
done
echo "=== C5 perceiver (graphed)" >> $L
timeout 120 python tools/bench_perceiver.py 2>&1 | tail -1 >> $L
cat $L | cut -c1-900
