#!/bin/bash
# One GPU call: per-launch durations and DRAM bytes of one serial-mode compression of config 2 (1 GiB), and an
# `ncu --set full` capture of every compression kernel on 256 MiB with its raw and per-source-line pages.
set -x
O=gpurun_out; T=${1:-r2x}; mkdir -p $O
ZSTDB200_SERIAL=1 timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:zb_ --csv --log-file $O/${T}_traffic_ncu.csv python tests/profile_one.py 1024 50 1 1 > $O/t_ncu.log 2>&1
python tools/ncu_traffic.py $O/${T}_traffic_ncu.csv | tee $O/${T}_traffic.txt
ZSTDB200_SERIAL=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"zb_walk|zb_parse|zb_merge|zb_literals|zb_sequences" -c 5 -o $O/${T}_all_256 -f python tests/profile_one.py 256 50 1 1 > $O/n1.log 2>&1
ncu -i $O/${T}_all_256.ncu-rep --page raw --csv > $O/${T}_ncu_full_all_256MiB.csv 2>/dev/null
for k in zb_merge zb_literals zb_sequences zb_parse zb_walk; do ncu -i $O/${T}_all_256.ncu-rep --page source --csv --kernel-name regex:$k > $O/${T}_src_$k.csv 2>/dev/null; done
ls -la $O | tail -15
