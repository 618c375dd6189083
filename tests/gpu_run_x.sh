set -x
cd $GRAFT_REPO_ROOT
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_all_v.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all_v.log)
(timeout 600 python bench.py --workload ada --steps 10 --warmup 3 > gpurun_out/bench_ada_v2.json 2> gpurun_out/bench_ada_v2.err; echo "rc=$?" >> gpurun_out/bench_ada_v2.err)
(timeout 300 python tests/bench_resnet.py > gpurun_out/bench_resnet_v.log 2>&1; echo "rc=$?" >> gpurun_out/bench_resnet_v.log)
