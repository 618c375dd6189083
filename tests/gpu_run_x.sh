set -x
cd $GRAFT_REPO_ROOT
(BSED_SANITIZE_PRECISION=tf32x3 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tests/gpu_sanitize.py > gpurun_out/memcheck_x3.log 2>&1; echo "rc=$?" >> gpurun_out/memcheck_x3.log)
(BSED_SANITIZE_PRECISION=tf32 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tests/gpu_sanitize.py > gpurun_out/memcheck_tf32.log 2>&1; echo "rc=$?" >> gpurun_out/memcheck_tf32.log)
