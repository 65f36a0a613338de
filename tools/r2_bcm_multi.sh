#!/bin/bash
# Multi-GPU BCM checks on N GPUs of one box (run under `gpurun --gpus N`):
#  1. torchrun: cugp_b200.BCM with the exchange inside the library (NCCL communicator owned by libcugp) vs golden C4
#  2. no Python, no torch: the reference's OWN driver distributed_gp/distributed_ver1.cpp, compiled unchanged against
#     include/cugp_shim, started once per GPU with CUGP_RANK / CUGP_WORLD / CUGP_NCCL_ID_FILE
set -u
mkdir -p gpurun_out
NG=${1:-$(nvidia-smi -L | wc -l)}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29561 \
    tools/bcm_nccl_check.py > gpurun_out/r2_bcm_nccl_check_${NG}gpu.log 2>&1
echo "torchrun parity rc=$? ($(grep -c OK gpurun_out/r2_bcm_nccl_check_${NG}gpu.log) OK lines)"; grep "rank 0" gpurun_out/r2_bcm_nccl_check_${NG}gpu.log | head -3
DRV=oracle/_ref/drivers/distributed_ver1
if [ -x $DRV ]; then
  rm -f /tmp/cugp_id_$$
  for r in $(seq 0 $((NG - 1))); do
    ( cd oracle/_ref/drivers/run_dist && CUGP_RANK=$r CUGP_WORLD=$NG CUGP_NCCL_ID_FILE=/tmp/cugp_id_$$ timeout 300 ../distributed_ver1 > ../../../../gpurun_out/r2_dist_ver1_rank${r}of${NG}.log 2>&1 ) &
  done
  wait
  for r in $(seq 0 $((NG - 1))); do echo "rank $r: $(grep -i "hyper\|final\|optim" gpurun_out/r2_dist_ver1_rank${r}of${NG}.log | tail -2 | tr '\n' ' ')"; done
else
  echo "no $DRV (built by tests/test_shim.py where /root/reference exists)"
fi
