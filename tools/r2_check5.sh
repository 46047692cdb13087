set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2z6_gputests.txt
timeout 300 python tests/bench_decode.py 2>&1 | tee gpurun_out/r2z6_decode.txt
