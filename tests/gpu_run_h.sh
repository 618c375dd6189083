set -x
cd $GRAFT_REPO_ROOT
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_all_h.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all_h.log)
(timeout 600 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_train.py tests/test_gpu_fpn.py tests/test_gpu_isp.py -q -s -k "eight or fixture or isp_step" > gpurun_out/gpu_new_h.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_new_h.log)
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "rc=$?" >> gpurun_out/bench_r2h.err)
(timeout 600 python bench.py --workload ada --steps 10 --warmup 3 > gpurun_out/bench_ada_h.json 2> gpurun_out/bench_ada_h.err; echo "rc=$?" >> gpurun_out/bench_ada_h.err)
(timeout 600 python bench.py --workload pseudo_label > gpurun_out/bench_pl_h.json 2> gpurun_out/bench_pl_h.err; echo "rc=$?" >> gpurun_out/bench_pl_h.err)
(timeout 600 python tests/bench_resnet.py > gpurun_out/bench_resnet_h.log 2>&1; echo "rc=$?" >> gpurun_out/bench_resnet_h.log)
du -sh gpurun_out
