set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 3000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/b_ncu.log 2>&1; tail -3 gpurun_out/r2a_launches.csv
ZSTDB200_SERIAL=1 timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:zb_ --csv --log-file gpurun_out/r2a_traffic.csv python tests/profile_one.py 1024 50 1 1 > gpurun_out/t_ncu.log 2>&1; tail -12 gpurun_out/r2a_traffic.csv
timeout 500 python tests/bench_configs.py c2 c3 c4 c5 2>&1 | tee gpurun_out/r2a_configs.txt | tail -8
