set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_all_f.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all_f.log)
(timeout 300 python tests/graph_probe.py > gpurun_out/graph_probe_f.log 2>&1; echo "rc=$?" >> gpurun_out/graph_probe_f.log)
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "rc=$?" >> gpurun_out/bench_r2f.err)
python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/plain_f.log 2>&1 && \
BSED_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02f_x3.csv python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/ncu_f1.log 2>&1
du -sh gpurun_out
