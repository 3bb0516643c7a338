#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 "$@" > gpurun_out/bench_n${N}_${name}.json 2> gpurun_out/bench_n${N}_${name}.err
  echo "== $name rc=$?"; tail -c 1500 gpurun_out/bench_n${N}_${name}.json; tail -2 gpurun_out/bench_n${N}_${name}.err
}
run grid_p2p --exchange p2p --workload grid_kuramoto_1e6
run cfg2_p2p --exchange p2p
run cfg5s_p2p --exchange p2p --workload cfg5_kuramoto_er_5e6
