#!/bin/bash
# 1-GPU job (final build): ncu --set full of K2b pass 1 and the cooperative apply at N3, after the plain run exits 0
python tools/select_once.py 675129632 > gpurun_out/sel_plain.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/sel_plain.log; exit 1; }
tail -1 gpurun_out/sel_plain.log
timeout 105 ncu --set full --clock-control none --import-source on -k "regex:select_hist1_kernel|select_apply_fused_kernel" -c 2 \
  -o gpurun_out/r2_prof_select_final -f python tools/select_once.py 675129632 > gpurun_out/ncu_sel_final.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_sel_final.log | cut -c1-200
