set -x
cd $GRAFT_REPO_ROOT
python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/plain_q.log 2>&1 && \
BSED_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02q_x3.csv python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/ncu_q1.log 2>&1
du -sh gpurun_out
