set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2z5_gputests.txt
timeout 300 python tests/variant_sweep.py c2 default 2>&1 | tee gpurun_out/r2z5_variants_c2.txt
timeout 300 python tests/variant_sweep.py c4 default 2>&1 | tee gpurun_out/r2z5_variants_c4.txt
timeout 300 python tests/bench_configs.py c3 c5 2>&1 | tee gpurun_out/r2z5_configs.txt
