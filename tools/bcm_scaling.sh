#!/bin/bash
# C4 (BCM 16 x 1500, 10000 test points) on 1, 2, 4, 8 GPUs of one box + the golden parity check at the widest size.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in ${NS:-1 2 4 8}; do
  [ $N -gt $NG ] && continue
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --workload c4 --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c4_${N}gpu.log 2> gpurun_out/bench_c4_${N}gpu.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --workload c4 --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c4_${N}gpu.log 2> gpurun_out/bench_c4_${N}gpu.err
  fi
  echo "N=$N rc=$?"; tail -1 gpurun_out/bench_c4_${N}gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('  step ms', d['ms_per_step'], 'pts/s', d['value'], 'phases', d.get('phases'))"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29555 tools/bcm_nccl_check.py > gpurun_out/bcm_nccl_check_${NG}gpu.log 2>&1; echo "parity rc=$?"; grep -c OK gpurun_out/bcm_nccl_check_${NG}gpu.log
# the exchange over peer memory against ncclAllReduce on the same handles, and the reference's own driver on two GPUs
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29556 tools/r2_peer_ab.py > gpurun_out/peer_ab_${NG}gpu.log 2>&1; echo "peer_ab rc=$?"; grep "exchange=\|relerr" gpurun_out/peer_ab_${NG}gpu.log
timeout 600 python -m pytest tests/test_shim.py -m gpu -x -q -k two_gpus > gpurun_out/shim_2gpu.log 2>&1; echo "shim 2gpu rc=$?"; tail -2 gpurun_out/shim_2gpu.log
