timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu25.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/pytest_gpu25.log
timeout 400 python tools/tpc_sweep.py 1,4,8 40000,100000 > gpurun_out/tpc_sweep2.txt 2>&1; cat gpurun_out/tpc_sweep2.txt
ONLY="trailing" bash tools/ncu_round.sh
python tools/ncu_summary.py gpurun_out/ncu_trailing_raw.csv | grep -E "duration|grid|DRAM read  |DRAM write  |tensor pipe" | paste - - - - - | sort -k2 -n | tail -2
