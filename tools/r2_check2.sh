set -x
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2z3_gpudecode.txt
timeout 300 python tests/bench_decode.py 2>&1 | tee gpurun_out/r2z3_decode.txt
timeout 300 python tests/bench_decode.py 256 2>&1 | tee gpurun_out/r2z3_decode256.txt
