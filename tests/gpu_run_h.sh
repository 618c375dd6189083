set -x
cd $GRAFT_REPO_ROOT
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_all_r.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all_r.log)
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2r.json 2> gpurun_out/bench_r2r.err; echo "rc=$?" >> gpurun_out/bench_r2r.err)
(timeout 600 python bench.py --workload ada --steps 10 --warmup 3 > gpurun_out/bench_ada_r.json 2> gpurun_out/bench_ada_r.err; echo "rc=$?" >> gpurun_out/bench_ada_r.err)
(timeout 600 python bench.py --workload pseudo_label > gpurun_out/bench_pl_r.json 2> gpurun_out/bench_pl_r.err; echo "rc=$?" >> gpurun_out/bench_pl_r.err)
(timeout 600 python bench.py --model crnn_fpn --steps 10 --warmup 3 > gpurun_out/bench_fpn_r.json 2> gpurun_out/bench_fpn_r.err; echo "rc=$?" >> gpurun_out/bench_fpn_r.err)
(timeout 600 python tests/bench_resnet.py > gpurun_out/bench_resnet_r.log 2>&1; echo "rc=$?" >> gpurun_out/bench_resnet_r.log)
du -sh gpurun_out
