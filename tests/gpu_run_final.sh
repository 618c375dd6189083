# Round-end single-GPU run: tests, every bench workload, launch list, per-launch tensor-core metrics, one --set full capture.
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q > $O/gpu_all_z.log 2>&1; echo "rc=$?" >> $O/gpu_all_z.log)
(timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_z.json 2> $O/bench_z.err; echo "rc=$?" >> $O/bench_z.err)
(timeout 600 python bench.py --workload ada --steps 10 --warmup 3 > $O/bench_ada_z.json 2> $O/bench_ada_z.err; echo "rc=$?" >> $O/bench_ada_z.err)
(timeout 600 python bench.py --workload pseudo_label > $O/bench_pl_z.json 2> $O/bench_pl_z.err; echo "rc=$?" >> $O/bench_pl_z.err)
(timeout 600 python bench.py --model crnn_fpn --steps 10 --warmup 3 > $O/bench_fpn_z.json 2> $O/bench_fpn_z.err; echo "rc=$?" >> $O/bench_fpn_z.err)
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_z.json 2> $O/bench_ref_z.err; echo "rc=$?" >> $O/bench_ref_z.err)
(timeout 600 python tests/bench_resnet.py > $O/bench_resnet_z.log 2>&1; echo "rc=$?" >> $O/bench_resnet_z.log)
(timeout 300 python tests/bench_conv.py tf32x3 > $O/bench_conv_z.log 2>&1; timeout 300 python tests/bench_conv.py tf32 >> $O/bench_conv_z.log 2>&1; timeout 300 python tests/bench_gemm.py tf32x3 >> $O/bench_conv_z.log 2>&1)
python tests/prof_step.py --steps 1 --warmup 3 > $O/plain_z.log 2>&1 && \
BSED_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_z.csv python tests/prof_step.py --steps 1 --warmup 3 > $O/ncu_z1.log 2>&1
BSED_GRAPH=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:tc_ -s 219 -c 73 --csv --log-file $O/tc_metrics_z.csv python tests/prof_step.py --steps 1 --warmup 3 > $O/ncu_z2.log 2>&1
BSED_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:tc_conv_col_kernel -s 55 -c 1 -o /tmp/conv_col_full python tests/prof_step.py --steps 1 --warmup 3 > $O/ncu_z3.log 2>&1
ncu -i /tmp/conv_col_full.ncu-rep --page raw --csv > $O/conv_col_full_z.csv 2>/dev/null
ncu -i /tmp/conv_col_full.ncu-rep --page details --csv > $O/conv_col_details_z.csv 2>/dev/null
du -sh $O
