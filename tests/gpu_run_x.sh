set -x
cd $GRAFT_REPO_ROOT
for i in 1 2 3; do
(timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/gpu_flaky_$i.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_flaky_$i.log)
done
