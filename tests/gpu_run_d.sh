set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_x3.py tests/test_gpu_ada_step.py tests/test_gpu_trainer_state.py -q -s > gpurun_out/x3_all_d.log 2>&1; echo "rc=$?" >> gpurun_out/x3_all_d.log)
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_all_d.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all_d.log)
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "rc=$?" >> gpurun_out/bench_r2d.err)
(timeout 600 python bench.py --workload ada --steps 10 --warmup 3 > gpurun_out/bench_ada_d.json 2> gpurun_out/bench_ada_d.err; echo "rc=$?" >> gpurun_out/bench_ada_d.err)
python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/plain_d.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02d_x3.csv python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/ncu_d1.log 2>&1
du -sh gpurun_out
