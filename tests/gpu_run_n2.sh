set -x
cd $GRAFT_REPO_ROOT
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_all_t.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all_t.log)
(timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; echo "rc=$?" >> gpurun_out/bench_t.err)
(BSED_WGRAD_ASIDE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_t_aside.json 2> gpurun_out/bench_t_aside.err; echo "rc=$?" >> gpurun_out/bench_t_aside.err)
