set -x
cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_gpu_da.py tests/test_gpu_ada_step.py tests/test_gpu_resnet.py -q > gpurun_out/gpu_disc_v.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_disc_v.log)
(timeout 600 python bench.py --workload ada --steps 10 --warmup 3 > gpurun_out/bench_ada_v.json 2> gpurun_out/bench_ada_v.err; echo "rc=$?" >> gpurun_out/bench_ada_v.err)
