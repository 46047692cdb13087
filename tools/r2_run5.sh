set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2e_gputests.txt
timeout 400 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2e_gpudecode.txt
timeout 400 python tests/variant_sweep.py c2 default p8 p8b5 2>&1 | tee gpurun_out/r2e_variants_c2.txt
timeout 500 python tests/variant_sweep.py c4 default df28 df28p4 df16p4 2>&1 | tee gpurun_out/r2e_variants_c4.txt
timeout 300 python tests/bench_decode.py 2>&1 | tee gpurun_out/r2e_decode.txt
