set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_x3.py tests/test_gpu_ada_step.py tests/test_gpu_pseudo_labeling.py -q -s > gpurun_out/x3_all_c.log 2>&1; echo "rc=$?" >> gpurun_out/x3_all_c.log)
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_all_c.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all_c.log)
(timeout 600 python bench.py --workload ada --steps 10 --warmup 3 > gpurun_out/bench_ada_c.json 2> gpurun_out/bench_ada_c.err; echo "rc=$?" >> gpurun_out/bench_ada_c.err)
(timeout 600 python bench.py --workload pseudo_label --steps 8 --warmup 1 > gpurun_out/bench_pl_c.json 2> gpurun_out/bench_pl_c.err; echo "rc=$?" >> gpurun_out/bench_pl_c.err)
(timeout 600 python bench.py --workload pseudo_label --steps 4 --warmup 1 --replicate 16 > gpurun_out/bench_pl16_c.json 2> gpurun_out/bench_pl16_c.err; echo "rc=$?" >> gpurun_out/bench_pl16_c.err)
python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/plain_c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02c_x3.csv python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/ncu_c1.log 2>&1
python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:tc_ -s 282 -c 94 --csv --log-file gpurun_out/tc_metrics_r02c_x3.csv python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/ncu_c2.log 2>&1
python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none -k regex:tc_conv_col_kernel -s 40 -c 2 -o /tmp/prof_col python tests/prof_step.py --steps 1 --warmup 3 > gpurun_out/ncu_c3.log 2>&1
ncu -i /tmp/prof_col.ncu-rep --page raw --csv > gpurun_out/conv_col_x3_full_r02c.csv 2>/dev/null
ls -la /tmp/prof_col.ncu-rep
du -sh gpurun_out
