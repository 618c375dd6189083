set -x
cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_gpu_x3.py tests/test_gpu_kernels.py tests/test_gpu_tf32.py -q -x > gpurun_out/gpu_x3_o.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_x3_o.log)
L=gpurun_out/bench_conv_o.log
: > $L
timeout 300 python tests/bench_conv.py tf32x3 >> $L 2>&1
timeout 300 python tests/bench_conv.py tf32 >> $L 2>&1
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err; echo "rc=$?" >> gpurun_out/bench_r2o.err)
du -sh gpurun_out
