#!/bin/bash
# multi-GPU round: N = number of visible GPUs.  Parity tests of the partitioned RHS, then bench lines.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/smi_multi.txt 2>&1
if [ "$N" -le 4 ]; then
  ( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_zzz_new_features.py tests/test_zzz_multi_stateful.py -m gpu -q -k "two_gpu" ) > gpurun_out/pytest_multi.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
  tail -8 gpurun_out/pytest_multi.log
fi
run() {  # name, extra args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 "$@" > gpurun_out/bench_n${N}_${name}.json 2> gpurun_out/bench_n${N}_${name}.err
  echo "== $name rc=$?"; tail -c 1800 gpurun_out/bench_n${N}_${name}.json; tail -3 gpurun_out/bench_n${N}_${name}.err
}
run cfg2_p2p --exchange p2p
run cfg2_p2p_pack --exchange p2p --pack
run cfg2_nccl --exchange nccl
run cfg5s_p2p --exchange p2p --workload cfg5_kuramoto_er_5e6
run cfg5s_p2p_pack --exchange p2p --workload cfg5_kuramoto_er_5e6 --pack
run grid_p2p --exchange p2p --workload grid_kuramoto_1e6
ls -la gpurun_out | head -40
