set -x
cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_gpu_x3.py tests/test_gpu_kernels.py tests/test_gpu_tf32.py -q -x > gpurun_out/gpu_x3_i.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_x3_i.log)
L=gpurun_out/bench_conv_i.log
: > $L
timeout 300 python tests/bench_conv.py tf32x3 >> $L 2>&1
BSED_COL_BSTAGES=3 timeout 300 python tests/bench_conv.py tf32x3 >> $L 2>&1
BSED_COL_BSTAGES=2 timeout 300 python tests/bench_conv.py tf32x3 >> $L 2>&1
BSED_TC_DEBUG=2 timeout 300 python tests/bench_conv.py tf32x3 >> $L 2>&1
BSED_TC_DEBUG=3 timeout 300 python tests/bench_conv.py tf32x3 >> $L 2>&1
BSED_TC_DEBUG=1 timeout 300 python tests/bench_conv.py tf32x3 >> $L 2>&1
timeout 300 python tests/bench_conv.py tf32 >> $L 2>&1
BSED_TC_DEBUG=2 timeout 300 python tests/bench_conv.py tf32 >> $L 2>&1
BSED_COL_BSTAGES=3 timeout 300 python tests/bench_conv.py tf32 >> $L 2>&1
(timeout 600 python bench.py --workload pseudo_label > gpurun_out/bench_pl_i.json 2> gpurun_out/bench_pl_i.err; echo "rc=$?" >> gpurun_out/bench_pl_i.err)
du -sh gpurun_out
