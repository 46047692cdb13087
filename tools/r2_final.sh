#!/bin/bash
# The gpurun command the final round-2 evidence under profiles/ (r2u_*) comes from (one B200), most important first:
#   gpurun --timeout 420 -- 'bash tools/r2_final.sh'
set -x
P=gpurun_out; mkdir -p $P
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $P/r2u_gputests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $P/r2u_smoke.txt
timeout 240 python bench.py > $P/r2u_bench.json 2> $P/r2u_bench.err; cut -c1-1200 $P/r2u_bench.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $P/r2u_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-decode > $P/b_ncu.log 2>&1; tail -2 $P/r2u_launches_bench_steps2.csv | cut -c1-300
ZSTDB200_SERIAL=1 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:zb_ --csv --log-file $P/r2u_traffic_ncu.csv python tests/profile_one.py 1024 50 1 1 > $P/t_ncu.log 2>&1
python tools/ncu_traffic.py $P/r2u_traffic_ncu.csv | tee $P/r2u_traffic.txt
timeout 120 python bench.py --impl reference --steps 3 --warmup 2 > $P/r2u_bench_reference.json 2>/dev/null; cut -c1-300 $P/r2u_bench_reference.json
timeout 150 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu > $P/r2u_bench_c4.json 2> $P/r2u_bench_c4.err; cut -c1-400 $P/r2u_bench_c4.json
timeout 150 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu > $P/r2u_bench_c5.json 2> $P/r2u_bench_c5.err; cut -c1-400 $P/r2u_bench_c5.json
ZSTDB200_SERIAL=1 timeout 200 ncu --set full --import-source on --clock-control none -k regex:"zb_walk|zb_parse|zb_merge|zb_literals|zb_sequences" -c 5 -o $P/r2u_all_256 -f python tests/profile_one.py 256 50 1 1 > $P/n1.log 2>&1
ncu -i $P/r2u_all_256.ncu-rep --page raw --csv > $P/r2u_ncu_full_all_256MiB.csv 2>/dev/null
timeout 150 python bench.py --config 3 --scale 0.25 --steps 5 --warmup 3 --no-cpu > $P/r2u_bench_c3_quarter.json 2> $P/r2u_bench_c3.err; cut -c1-400 $P/r2u_bench_c3_quarter.json
timeout 120 python tests/wave_sweep.py c2 3 2>&1 | grep -v Warning | tee $P/r2u_wave_sweep.txt
