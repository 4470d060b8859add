set -e
python tools/run_once.py C3 2 > gpurun_out/plain_C3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tet_march|tet_first" -c 3 -o gpurun_out/r01b_full_C3 -f python tools/run_once.py C3 1 > gpurun_out/ncu_full_C3b.log 2>&1
python tools/ncu_summary.py gpurun_out/r01b_full_C3.ncu-rep gpurun_out/r01b_ncu_full_C3.csv
