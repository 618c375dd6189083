set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_trainer_state.py tests/test_gpu_ada_step.py tests/test_gpu_train.py tests/test_gpu_fpn.py -q -s > gpurun_out/t_e.log 2>&1; echo "rc=$?" >> gpurun_out/t_e.log)
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_all_e.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all_e.log)
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "rc=$?" >> gpurun_out/bench_r2e.err)
(timeout 600 python tests/bench_resnet.py > gpurun_out/bench_resnet_e.log 2>&1; echo "rc=$?" >> gpurun_out/bench_resnet_e.log)
du -sh gpurun_out
