#!/bin/bash
# Multi-GPU evidence of one round, run on an 8-GPU B200 box (gpurun --gpus 8): the 4- and 8-GPU bit-identity tests
# (NCCL gather and peer-memory gather), the bench line at N = 8 (C2 replicas + parallel tempering C3 / C4 records with
# the bit-identity flag) and parallel tempering C3 / C4 with the three gather modes.  Outputs under gpurun_out/.
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p $out

timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "4-even or 8-even or 8-uneven" > $out/${tag}_pytest_multi_8gpu.log 2>&1
tail -4 $out/${tag}_pytest_multi_8gpu.log
run() { # name, env, args...
  name=$1; shift; envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 "$@" > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err
  tail -c 1500 $out/${tag}_${name}.json; echo
}
run bench_c2_8gpu "CSMC_PEER_GATHER=0" --steps 5 --warmup 3
for mode in 0 1 2; do
  run bench_pt_c3_8gpu_gather$mode "CSMC_PEER_GATHER=$mode" --workload C3 --steps 5 --warmup 2
done
for mode in 0 2; do
  run bench_pt_c4_8gpu_gather$mode "CSMC_PEER_GATHER=$mode" --workload C4 --steps 5 --warmup 2
done
