set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_x3.py tests/test_gpu_tf32.py tests/test_gpu_train.py tests/test_gpu_fpn.py tests/test_gpu_kernels.py -q -x > gpurun_out/gpu_ring.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_ring.log)
(timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ring.json 2> gpurun_out/bench_ring.err; echo "rc=$?" >> gpurun_out/bench_ring.err)
(BSED_NO_OP_RING=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_noring.json 2> gpurun_out/bench_noring.err; echo "rc=$?" >> gpurun_out/bench_noring.err)
