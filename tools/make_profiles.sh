#!/bin/bash
# Regenerates the round's measurement artefacts on a B200 box (run through gpurun):
#   gpurun --timeout 1800 -- 'bash tools/make_profiles.sh r01c'
# Outputs land in gpurun_out/ (copied into profiles/ by hand after review).
set -x
tag=${1:-r01c}
out=gpurun_out
mkdir -p $out
python tools/time_compare.py C1 C2 C3 C5 C4:8 > $out/${tag}_time_compare.txt 2>&1
python tools/stage_times.py C1 C2 C3 C5 C4:8 > $out/${tag}_stage_times.txt 2>&1
python bench.py --steps 50 --warmup 5 2>/dev/null | tail -1 > $out/${tag}_bench_ours_C2.json
python bench.py --impl reference --steps 50 --warmup 5 2>/dev/null | tail -1 > $out/${tag}_bench_reference_C2.json
# launch list of the bench command itself (its timed numbers above come from the runs WITHOUT ncu)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_bench_C2.csv python bench.py --steps 3 --warmup 3 > $out/ncu_bench.log 2>&1
for cfg in C2 C3 C5; do
  python tools/run_once.py $cfg 2 > $out/plain_$cfg.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_$cfg.csv python tools/run_once.py $cfg 2 > $out/ncu_$cfg.log 2>&1
done
python tools/run_once.py C4 2 8 > $out/plain_C4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_C4x8.csv python tools/run_once.py C4 2 8 > $out/ncu_C4.log 2>&1
python tools/launch_summary.py $out/${tag}_launches_bench_C2.csv $out/${tag}_launches_C2.csv $out/${tag}_launches_C3.csv $out/${tag}_launches_C5.csv $out/${tag}_launches_C4x8.csv > $out/${tag}_launch_shares.txt 2>&1
# full captures of the dominant kernels (one launch each)
ncu --set full --clock-control none --import-source on -k regex:"tri_render" -c 2 -o $out/${tag}_full_C2 -f python tools/run_once.py C2 1 > $out/ncu_full_C2.log 2>&1
python tools/ncu_summary.py $out/${tag}_full_C2.ncu-rep $out/${tag}_ncu_full_C2.csv
ncu --set full --clock-control none --import-source on -k regex:"tet_march|tet_first" -c 4 -o $out/${tag}_full_C3 -f python tools/run_once.py C3 1 > $out/ncu_full_C3.log 2>&1
python tools/ncu_summary.py $out/${tag}_full_C3.ncu-rep $out/${tag}_ncu_full_C3.csv
ncu --set full --clock-control none --import-source on -k regex:"rs_onesweep|duplicate|tile_ranges|preprocess|inclusive_scan" -c 12 -o $out/${tag}_full_bin_C5 -f python tools/run_once.py C5 1 > $out/ncu_full_bin_C5.log 2>&1
python tools/ncu_summary.py $out/${tag}_full_bin_C5.ncu-rep $out/${tag}_ncu_full_bin_C5.csv
echo done
