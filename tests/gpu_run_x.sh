set -x
cd $GRAFT_REPO_ROOT
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_all_w.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all_w.log)
(timeout 600 python bench.py --workload ada --steps 10 --warmup 3 > gpurun_out/bench_ada_w.json 2> gpurun_out/bench_ada_w.err; echo "rc=$?" >> gpurun_out/bench_ada_w.err)
(timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke_w.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_w.log)
