# The gpurun command the committed round-2 evidence under profiles/ comes from (one B200):
#   gpurun --timeout 2400 -- 'bash tools/r2_evidence.sh'
set -x
P=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $P/r2z_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $P/r2z_smoke.txt
timeout 400 python bench.py > $P/r2z_bench.json 2> $P/r2z_bench.err; tail -c 3500 $P/r2z_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 2 > $P/r2z_bench_reference.json 2>/dev/null; cut -c1-400 $P/r2z_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $P/r2z_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-decode > $P/b_ncu.log 2>&1; tail -2 $P/r2z_launches_bench_steps2.csv | cut -c1-300
ZSTDB200_SERIAL=1 timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:zb_ --csv --log-file $P/r2z_traffic_ncu.csv python tests/profile_one.py 1024 50 1 1 > $P/t_ncu.log 2>&1
python tools/ncu_traffic.py $P/r2z_traffic_ncu.csv | tee $P/r2z_traffic.txt
ZSTDB200_SERIAL=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"zb_walk|zb_parse" -c 2 -o $P/r2z_match_256 -f python tests/profile_one.py 256 50 1 1 > $P/n1.log 2>&1
timeout 500 python tests/bench_configs.py c2 c3 c4 c5 2>&1 | tee $P/r2z_configs.txt | tail -6
for c in 3 4 5; do timeout 400 python bench.py --config $c --steps 5 --warmup 3 > $P/r2z_bench_c$c.json 2> $P/r2z_bench_c$c.err; cut -c1-700 $P/r2z_bench_c$c.json; done
timeout 300 python tests/bench_decode.py 2>&1 | tee $P/r2z_decode.txt
timeout 300 python tests/bench_decode.py 256 2>&1 | tee $P/r2z_decode256.txt
timeout 600 compute-sanitizer --tool memcheck python tests/sanitize_small.py 2>&1 | tail -6 | tee $P/r2z_sanitizer_memcheck.txt
