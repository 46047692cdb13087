set -x
export ZSTDB200_SERIAL=1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"zb_walk|zb_parse|zb_merge" -c 3 -o gpurun_out/r2b_match_256 -f python tests/profile_one.py 256 50 1 1 > gpurun_out/n1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"zb_literals|zb_sequences" -c 2 -o gpurun_out/r2b_entropy_256 -f python tests/profile_one.py 256 50 1 1 > gpurun_out/n2.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:zb_ --csv --log-file gpurun_out/r2b_c4_traffic.csv python tests/profile_one.py 1024 90 3 1 > gpurun_out/n3.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:zb_ --csv --log-file gpurun_out/r2b_c3_traffic.csv python tests/profile_one.py 1024 30 -3 1 > gpurun_out/n4.log 2>&1
ls -la gpurun_out
