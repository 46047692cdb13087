set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2f_gputests.txt
timeout 400 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2f_gpudecode.txt
timeout 300 python tests/bench_decode.py 2>&1 | tee gpurun_out/r2f_decode.txt
timeout 300 python tests/bench_decode.py 256 2>&1 | tee gpurun_out/r2f_decode256.txt
timeout 500 python tests/bench_configs.py c2 c3 c4 c5 2>&1 | tee gpurun_out/r2f_configs.txt | tail -8
