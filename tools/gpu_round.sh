#!/bin/bash
# One GPU-box session: parity tests, smoke, default bench, ncu launch list of the bench command, one full capture of the top kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke_rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py --steps ${STEPS:-3} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench_rc=$?"
tail -1 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
if [ "${REF:-1}" = "1" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref_rc=$?"; tail -1 gpurun_out/bench_ref.log
fi
if [ "${NCU:-1}" = "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-3000} --csv --log-file gpurun_out/launches_bench.csv \
   python bench.py --steps 1 --warmup 3 --no-extra ${BENCH_ARGS:-} > gpurun_out/ncu_bench.log 2>&1; echo "ncu_list_rc=$?"
fi
