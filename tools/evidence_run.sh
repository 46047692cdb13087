set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; python -c "import __graft_entry__ as g; g.smoke()"; timeout 200 python tests/variant_sweep.py c2 default pfnone 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -c 2500 gpurun_out/bench_r1_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/b_ncu.log 2>&1; tail -3 gpurun_out/r1_launches_final.csv
ZSTDB200_SERIAL=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:zb_ --csv --log-file gpurun_out/r1_traffic.csv python tests/profile_one.py 1024 50 1 1 > gpurun_out/t_ncu.log 2>&1; cat gpurun_out/r1_traffic.csv | tail -9
timeout 500 python tests/bench_configs.py c2 c3 c4 c5 c5full 2>&1 | tee gpurun_out/configs_r1_final.txt | tail -6
timeout 200 python tests/latency_sweep.py default 2>&1 | tee gpurun_out/latency_r1_final.txt | tail -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/bench_r1_reference.json 2>/dev/null; cut -c1-600 gpurun_out/bench_r1_reference.json
