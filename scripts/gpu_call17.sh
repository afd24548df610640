#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for wl in c3 c2 c5; do
python scripts/ncu_one.py $wl 64 > gpurun_out/c17_plain_$wl.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_render_mega|k_primary" -s 4 -c 2 --csv --log-file gpurun_out/c17_launches_$wl.csv python scripts/ncu_one.py $wl 64 > gpurun_out/c17_ncu_$wl.log 2>&1
python scripts/ncu_one.py $wl 64 3 16 > gpurun_out/c17_plain_inline_$wl.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_render_mega|k_primary" -s 2 -c 1 --csv --log-file gpurun_out/c17_launches_inline_$wl.csv python scripts/ncu_one.py $wl 64 3 16 > gpurun_out/c17_ncu_inline_$wl.log 2>&1
done
cat gpurun_out/c17_plain_*.log; grep -h "k_render\|k_primary" gpurun_out/c17_launches_*.csv | cut -d, -f5,13,15 | head -60
