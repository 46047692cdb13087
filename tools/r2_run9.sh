set -x
timeout 400 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2h_gpudecode.txt
timeout 300 python tests/bench_decode.py 2>&1 | tee gpurun_out/r2h_decode.txt
bash tools/r2_evidence.sh
