set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_all_g.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all_g.log)
(timeout 300 python tests/graph_probe.py > gpurun_out/graph_probe_g.log 2>&1; echo "rc=$?" >> gpurun_out/graph_probe_g.log)
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "rc=$?" >> gpurun_out/bench_r2g.err)
(BSED_WGRAD_ASIDE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2g_aside.json 2> gpurun_out/bench_r2g_aside.err; echo "rc=$?" >> gpurun_out/bench_r2g_aside.err)
python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/plain_g.log 2>&1 && \
BSED_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02g_x3.csv python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/ncu_g1.log 2>&1
du -sh gpurun_out
