set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
(timeout 600 python -m pytest tests/test_gpu_x3.py -x -q -s -k "gemm_nt_3xtf32 or conv3x3_3xtf32 or single_pass" > gpurun_out/x3_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/x3_kernels.log)
(BSED_X3_INPLACE=1 timeout 600 python -m pytest tests/test_gpu_x3.py -q -s -k "gemm_nt_3xtf32 or conv3x3_3xtf32 or single_pass" > gpurun_out/x3_kernels_inplace.log 2>&1; echo "rc=$?" >> gpurun_out/x3_kernels_inplace.log)
(timeout 900 python -m pytest tests/test_gpu_x3.py -q -s > gpurun_out/x3_all.log 2>&1; echo "rc=$?" >> gpurun_out/x3_all.log)
(timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all.log)
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "rc=$?" >> gpurun_out/bench_r2a.err)
tail -5 gpurun_out/x3_kernels.log gpurun_out/x3_all.log gpurun_out/gpu_all.log gpurun_out/bench_r2a.err
