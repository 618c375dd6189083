# usage: bash tests/gpu_run_n_short.sh N tag      (the two training benches only)
N=$1; TAG=$2
set -x
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err; echo "rc=$?" >> gpurun_out/bench_n${N}_${TAG}.err)
(timeout 300 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --workload ada > gpurun_out/bench_ada_n${N}_${TAG}.json 2> gpurun_out/bench_ada_n${N}_${TAG}.err; echo "rc=$?" >> gpurun_out/bench_ada_n${N}_${TAG}.err)
if [ "$N" = "2" ]; then
(timeout 300 python -m pytest tests/test_gpu_dp.py -q > gpurun_out/pytest_dp_n2_${TAG}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dp_n2_${TAG}.log)
fi
true
