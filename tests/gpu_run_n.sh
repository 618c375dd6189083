# usage: bash tests/gpu_run_n.sh N tag
N=$1; TAG=$2
set -x
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 400 $TR --master-port 29512 tests/dp_train_check.py > gpurun_out/dp_train_check_n${N}_${TAG}.log 2>&1; echo "rc=$?" >> gpurun_out/dp_train_check_n${N}_${TAG}.log)
(timeout 300 $TR --master-port 29511 tests/dp_fused_check.py > gpurun_out/dp_fused_check_n${N}_${TAG}.log 2>&1; echo "rc=$?" >> gpurun_out/dp_fused_check_n${N}_${TAG}.log)
(timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err; echo "rc=$?" >> gpurun_out/bench_n${N}_${TAG}.err)
(timeout 400 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --workload ada > gpurun_out/bench_ada_n${N}_${TAG}.json 2> gpurun_out/bench_ada_n${N}_${TAG}.err; echo "rc=$?" >> gpurun_out/bench_ada_n${N}_${TAG}.err)
if [ "$N" = "8" ]; then
(timeout 300 $TR --master-port 29515 bench.py --gpus $N --steps 8 --warmup 1 --workload pseudo_label > gpurun_out/bench_pl_n${N}_${TAG}.json 2> gpurun_out/bench_pl_n${N}_${TAG}.err; echo "rc=$?" >> gpurun_out/bench_pl_n${N}_${TAG}.err)
(timeout 300 $TR --master-port 29516 bench.py --gpus $N --steps 4 --warmup 1 --workload pseudo_label --replicate 64 > gpurun_out/bench_pl64_n${N}_${TAG}.json 2> gpurun_out/bench_pl64_n${N}_${TAG}.err; echo "rc=$?" >> gpurun_out/bench_pl64_n${N}_${TAG}.err)
fi
if [ "$N" = "2" ]; then
(timeout 600 python -m pytest tests/test_gpu_dp.py -q > gpurun_out/pytest_dp_n2_${TAG}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dp_n2_${TAG}.log)
fi
true
