#!/bin/bash
# End-of-round evidence on one B200: GPU test suite, the bench lines (C5 default, C2, C3, C4, the reference arm), smoke(),
# then the ncu launch list of the C2 line and one --set full capture of the fused block step (64-row tiles, n = 2048).
set -u
mkdir -p gpurun_out/ev
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/ev/pytest_gpu.txt
python bench.py > gpurun_out/ev/bench_c5.json 2> gpurun_out/ev/bench_c5.err; echo c5 rc=$?
for w in c2 c3 c4; do python bench.py --workload $w --steps 5 --warmup 3 --no-extra > gpurun_out/ev/bench_$w.json 2> gpurun_out/ev/bench_$w.err; echo $w rc=$?; done
python bench.py --impl reference > gpurun_out/ev/bench_ref_c5.json 2> gpurun_out/ev/bench_ref.err; echo ref rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --workload c2 --steps 1 --warmup 3 --no-extra --no-cpu > gpurun_out/ev/c2_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/ev/r2_launches_c2_final.csv \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-extra --no-cpu > gpurun_out/ev/ncu_c2.log 2>&1; echo "ncu_list_rc=$?"
timeout 100 python tools/run_eval.py 2048 2 > gpurun_out/ev/eval2048_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:chol_step_kernel" -s 18 -c 2 \
    -o gpurun_out/ev/r2_ncu_chol_step64 -f python tools/run_eval.py 2048 2 > gpurun_out/ev/ncu_step.log 2>&1; echo "ncu_step rc=$?"
ncu -i gpurun_out/ev/r2_ncu_chol_step64.ncu-rep --page raw --csv > gpurun_out/ev/r2_ncu_chol_step64_raw.csv 2>/dev/null
ls -la gpurun_out/ev | tail -20
