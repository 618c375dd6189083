set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_crnn.py -q -s > gpurun_out/gpu_resnet_u.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_resnet_u.log)
(timeout 600 python tests/bench_resnet.py > gpurun_out/bench_resnet_u.log 2>&1; echo "rc=$?" >> gpurun_out/bench_resnet_u.log)
