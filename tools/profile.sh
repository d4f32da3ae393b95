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
E="python bench_extra.py --only config3,config4 --out $O/prof_extra.json"
timeout 300 $E > $O/prof_plain_extra.log 2>&1 || { echo "plain bench_extra failed"; tail -3 $O/prof_plain_extra.log; exit 1; }
timeout 600 ncu --metrics $M --clock-control none -k regex:"sort_select|box_decode_selected|nms_kernel|decode_vec|sort_topk|roi_kernel|roi_big" -c 700 --csv --log-file $O/launches_config3_config4_r02.csv $E > $O/ncu_launches_extra.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sort_select_kernel|box_decode_selected_kernel|nms_kernel|roi_kernel|roi_big_kernel" -s 300 -c 12 -o $O/prof_r02_config3_config4 $E > $O/ncu_extra.log 2>&1
ls -la $O/*.ncu-rep | tail -3
