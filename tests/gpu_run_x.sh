set -x
cd $GRAFT_REPO_ROOT
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_all_y.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_all_y.log)
(timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke_y.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_y.log)
(timeout 900 python bench.py > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; echo "rc=$?" >> gpurun_out/bench_y.err)
