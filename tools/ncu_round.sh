#!/bin/bash
# ncu --set full captures of the kernels the north star names (one launch each), on commands that already ran
# clean without ncu.  Outputs: gpurun_out/ncu_<name>.ncu-rep + raw CSV pages.
set -u
mkdir -p gpurun_out
N=${N:-40000}
timeout 200 python tools/solve_timing.py $N > gpurun_out/ncu_plain_solve.log 2>&1 || { echo plain_failed; tail -3 gpurun_out/ncu_plain_solve.log; exit 1; }
ONLY=${ONLY:-}
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  if [ -n "$ONLY" ] && [[ " $ONLY " != *" $name "* ]]; then return; fi
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s $skip -c ${COUNT:-1} -o gpurun_out/ncu_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu_$name rc=$?"
  ncu -i gpurun_out/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_raw.csv 2>/dev/null
}
COUNT=8 cap trailing 'dgemm_ws_kernel<.int.128, .int.128' 6 python tools/chol_only.py $N 1
cap cov_lower 'cov_tile_kernel<.int.0' 0 python tools/chol_only.py $N 1
cap bwd_update 'trsv_bwd_update_kernel' 10 python tools/solve_timing.py $N
cap bwd_chain 'trsv_bwd_chain_kernel' 10 python tools/solve_timing.py $N
timeout 200 python tools/run_eval.py 16000 1 > gpurun_out/ncu_plain_eval.log 2>&1 || { echo plain_eval_failed; exit 1; }
cap grad_trace 'cov_tile_kernel<.int.3' 0 python tools/run_eval.py 16000 1
cap gemv_t 'gemv_t_kernel' 0 python tools/run_eval.py 16000 1
cap lauum 'dgemm_ws_kernel<.int.128, .int.128, .int.2, .int.4, .int.3, .bool.0, .bool.0' 0 python tools/run_eval.py 16000 1
cap cov_cross 'cov_tile_kernel<.int.2' 0 python tools/run_eval.py 16000 1
ls -la gpurun_out/ncu_*.ncu-rep
