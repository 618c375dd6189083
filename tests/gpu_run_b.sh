set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_x3.py -q -s > gpurun_out/x3_all_b.log 2>&1; echo "rc=$?" >> gpurun_out/x3_all_b.log)
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_all_b.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all_b.log)
python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02b_x3.csv python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/ncu_b1.log 2>&1
python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/plain_b2.log 2>&1 && \
ncu --set full --clock-control none -k regex:"tc_conv_col_kernel|tc_kmajor_kernel" -s 198 -c 66 -o gpurun_out/prof_tc_x3_r02b python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/ncu_b2.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/x3_all_b.log gpurun_out/gpu_all_b.log
