set -x
cd $GRAFT_REPO_ROOT
python tests/prof_step.py --ada --steps 1 --warmup 3 > gpurun_out/plain_ada.log 2>&1 && \
BSED_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_ada3.csv python tests/prof_step.py --ada --steps 1 --warmup 3 > gpurun_out/ncu_ada.log 2>&1
