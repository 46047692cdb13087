set -x
timeout 400 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2h_gpudecode.txt
timeout 300 python tests/bench_decode.py 2>&1 | tee gpurun_out/r2h_decode.txt
timeout 300 python tests/bench_decode.py 256 2>&1 | tee gpurun_out/r2h_decode256.txt
timeout 300 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -3
