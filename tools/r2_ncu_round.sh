#!/bin/bash
# Round-2 evidence run on one B200 (gpurun): the default bench line, the ncu launch list of the same command, and
# ncu --set full captures of the kernels the rooflines are quoted on.  Every ncu command ran clean without ncu first.
set -u
mkdir -p gpurun_out
STEPS=${STEPS:-3}
timeout 1500 python bench.py --steps $STEPS --warmup 3 > gpurun_out/r2_bench_c5.log 2> gpurun_out/r2_bench_c5.err; echo "bench_rc=$?"; tail -c 600 gpurun_out/r2_bench_c5.log; tail -2 gpurun_out/r2_bench_c5.err
timeout 600 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu > gpurun_out/r2_bench_plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-extra --no-cpu > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu_list_rc=$?"
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 regex=$2 skip=$3 count=$4; shift 4
  timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s $skip -c $count -o gpurun_out/r2_ncu_$name -f "$@" > gpurun_out/r2_ncu_$name.log 2>&1
  echo "ncu_$name rc=$?"
  ncu -i gpurun_out/r2_ncu_$name.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${name}_raw.csv 2>/dev/null
}
timeout 300 python tools/chol_only.py 100000 1 > gpurun_out/r2_chol_only_plain.log 2>&1 || { echo plain_failed; exit 1; }
# C5 itself: 80 GB of live device memory -- application replay (no per-kernel save / restore of 80 GB) and only the
# metrics the roofline line needs; the 10 first launches of the trailing-update template include U2(0), the longest
timeout 1500 ncu --replay-mode application --clock-control none --kernel-name-base demangled \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__grid_size \
    -k 'regex:dgemm_ws_kernel<.int.128, .int.128, .int.2, .int.4, .int.3, .bool.1, .bool.1' -c 10 -o gpurun_out/r2_ncu_trailing_c5 -f \
    python tools/chol_only.py 100000 1 > gpurun_out/r2_ncu_trailing_c5.log 2>&1; echo "ncu_trailing_c5 rc=$?"
ncu -i gpurun_out/r2_ncu_trailing_c5.ncu-rep --page raw --csv > gpurun_out/r2_ncu_trailing_c5_raw.csv 2>/dev/null
cap cov_lower 'cov_tile_kernel<.int.0' 0 1 python tools/chol_only.py 40000 1
timeout 100 python tools/run_eval.py 1500 2 > gpurun_out/r2_eval1500_plain.log 2>&1 || { echo plain_eval_failed; exit 1; }
cap chol_step 'chol_step_kernel' 14 2 python tools/run_eval.py 1500 2
cap grad_trace 'cov_tile_kernel<.int.3' 1 1 python tools/run_eval.py 1500 2
timeout 100 python tools/run_eval.py 16000 1 > gpurun_out/r2_eval16000_plain.log 2>&1 || { echo plain_eval_failed; exit 1; }
cap cov_cross 'cov_tile_kernel<.int.2' 0 1 python tools/run_eval.py 16000 1
ls -la gpurun_out/r2_ncu_*.ncu-rep
