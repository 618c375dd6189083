set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_x3.py tests/test_gpu_kernels.py tests/test_gpu_tf32.py -q -x > gpurun_out/gpu_cat.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_cat.log)
timeout 300 python tests/bench_conv.py tf32x3 > gpurun_out/bench_conv_cat.log 2>&1
(timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cat.json 2> gpurun_out/bench_cat.err; echo "rc=$?" >> gpurun_out/bench_cat.err)
