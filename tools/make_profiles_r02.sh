#!/bin/bash
# Regenerates round 2's measurement artefacts on a B200 box (run through gpurun):
#   gpurun --timeout 2400 -- 'bash tools/make_profiles_r02.sh r02b'
# Outputs land in gpurun_out/ (copied into profiles/ after review).  Numbers printed by runs under ncu are never
# bench values: every ncu pass is preceded by the same command without ncu.
tag=${1:-r02b}
out=gpurun_out
mkdir -p $out
python tools/time_compare.py C1 C2 C3 C5 C4:8 > $out/${tag}_time_compare.txt 2>&1
python tools/stage_times.py C1 C2 C3 C5 C4:8 > $out/${tag}_stage_times.txt 2>&1
tools/_bin/sort_vs_cub 32000000 16384 > $out/${tag}_sort_vs_cub.txt 2>&1
tools/_bin/sort_vs_cub 16800000 32768 >> $out/${tag}_sort_vs_cub.txt 2>&1
tools/_bin/sort_vs_cub 730000 4096 >> $out/${tag}_sort_vs_cub.txt 2>&1
# the two bench arms (default workload C4, N = 1)
python bench.py --steps 10 --warmup 3 2> $out/${tag}_bench_ours.err | tail -1 > $out/${tag}_bench_ours_C4.json
python bench.py --impl reference --steps 5 --warmup 3 2> $out/${tag}_bench_ref.err | tail -1 > $out/${tag}_bench_reference_C4.json
python bench.py --workload C2 --steps 50 --warmup 5 --no-per-config 2>/dev/null | tail -1 > $out/${tag}_bench_ours_C2.json
python bench.py --workload C2 --impl reference --steps 30 --warmup 5 --no-per-config 2>/dev/null | tail -1 > $out/${tag}_bench_reference_C2.json
# launch list of the bench command itself
python bench.py --steps 2 --warmup 3 --no-per-config > $out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches_bench_C4.csv python bench.py --steps 2 --warmup 3 --no-per-config > $out/ncu_bench.log 2>&1
for cfg in C2 C3 C5; do
  python tools/run_once.py $cfg 2 > $out/plain_$cfg.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_$cfg.csv python tools/run_once.py $cfg 2 > $out/ncu_$cfg.log 2>&1
done
python tools/run_once.py C4 2 8 > $out/plain_C4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_C4x8.csv python tools/run_once.py C4 2 8 > $out/ncu_C4.log 2>&1
python tools/launch_summary.py $out/${tag}_launches_bench_C4.csv $out/${tag}_launches_C2.csv $out/${tag}_launches_C3.csv $out/${tag}_launches_C5.csv $out/${tag}_launches_C4x8.csv > $out/${tag}_launch_shares.txt 2>&1
# full captures of the dominant kernels (one launch each)
ncu --set full --clock-control none --import-source on -k regex:"tri_render|tri_grad_finish" -c 3 -o $out/${tag}_full_C4x8 -f python tools/run_once.py C4 1 8 > $out/ncu_full_C4x8.log 2>&1
python tools/ncu_summary.py $out/${tag}_full_C4x8.ncu-rep $out/${tag}_ncu_full_C4x8.csv
ncu --set full --clock-control none --import-source on -k regex:"tri_render" -c 2 -o $out/${tag}_full_C2 -f python tools/run_once.py C2 1 > $out/ncu_full_C2.log 2>&1
python tools/ncu_summary.py $out/${tag}_full_C2.ncu-rep $out/${tag}_ncu_full_C2.csv
ncu --set full --clock-control none --import-source on -k regex:"tet_march|tet_first" -c 4 -o $out/${tag}_full_C3 -f python tools/run_once.py C3 1 > $out/ncu_full_C3.log 2>&1
python tools/ncu_summary.py $out/${tag}_full_C3.ncu-rep $out/${tag}_ncu_full_C3.csv
ncu --set full --clock-control none --import-source on -k regex:"rs_onesweep|duplicate|tile_ranges|preprocess|inclusive_scan|rs_hist" -c 14 -o $out/${tag}_full_bin_C5 -f python tools/run_once.py C5 1 > $out/ncu_full_bin_C5.log 2>&1
python tools/ncu_summary.py $out/${tag}_full_bin_C5.ncu-rep $out/${tag}_ncu_full_bin_C5.csv
echo done
