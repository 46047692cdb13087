set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2z2_gputests.txt
timeout 300 python tests/variant_sweep.py c2 default 2>&1 | tee gpurun_out/r2z2_variants_c2.txt
timeout 300 python tests/variant_sweep.py c4 default 2>&1 | tee gpurun_out/r2z2_variants_c4.txt
timeout 400 python tests/bench_configs.py c3 c5 c5full 2>&1 | tee gpurun_out/r2z2_configs.txt | tail -5
timeout 400 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2z2_bench_c5.json 2> gpurun_out/r2z2_bench_c5.err; cut -c1-500 gpurun_out/r2z2_bench_c5.json
