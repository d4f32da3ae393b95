#!/usr/bin/env bash
# ncu evidence for profiles/ (run under gpurun; every profiled command first exits 0 without ncu):
#   launch list of a short bench run (per-launch device time + DRAM bytes: compare SHARES, cold-cache + serialised),
#   `--set full` captures of the step's kernels, of the dense-regime chain (config 3) and of K5 on 4096 ROIs (config 4).
set -uo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 1 --reps 2 --e2e-reps 1 --no-extra --no-cpu-baseline --no-parity"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active"
timeout 300 $B > $O/prof_plain_bench.log 2>&1 || { echo "plain bench failed"; tail -3 $O/prof_plain_bench.log; exit 1; }
timeout 600 ncu --metrics $M --clock-control none -c 500 --csv --log-file $O/launches_r02.csv $B > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"letterbox_kernel|decode_vec_kernel|postprocess_small_kernel|roi_det_kernel" -s 40 -c 8 -o $O/prof_r02_step $B > $O/ncu_step.log 2>&1
E3="python bench_extra.py --only config3 --out $O/prof_extra3.json"
E4="python bench_extra.py --only config4 --out $O/prof_extra4.json"
timeout 300 $E3 > $O/prof_plain_extra3.log 2>&1 || { echo "plain bench_extra (config3) failed"; tail -3 $O/prof_plain_extra3.log; exit 1; }
timeout 300 $E4 > $O/prof_plain_extra4.log 2>&1 || { echo "plain bench_extra (config4) failed"; tail -3 $O/prof_plain_extra4.log; exit 1; }
timeout 600 ncu --metrics $M --clock-control none -k regex:"sort_select|box_decode_selected|nms_kernel|decode_vec|sort_topk" -c 300 --csv --log-file $O/launches_config3_r02.csv $E3 > $O/ncu_launches_extra3.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:"roi_kernel|roi_big" -c 40 --csv --log-file $O/launches_config4_r02.csv $E4 > $O/ncu_launches_extra4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"decode_vec_kernel|sort_select_kernel|box_decode_selected_kernel|nms_kernel" -s 8 -c 10 -o $O/prof_r02_config3 $E3 > $O/ncu_extra3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"roi_kernel|roi_big_kernel" -s 4 -c 4 -o $O/prof_r02_config4 $E4 > $O/ncu_extra4.log 2>&1
ls -la $O/*.ncu-rep | tail -3
