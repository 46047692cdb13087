set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2d_gputests.txt
timeout 300 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2d_gpudecode.txt
timeout 600 python tests/variant_sweep.py c2 default p2b3 p2b4 p1b2 2>&1 | tee gpurun_out/r2d_variants_c2.txt
ZSTDB200_DEBUG=1 timeout 300 python tests/variant_sweep.py c4 default 2>&1 | tail -5 | tee gpurun_out/r2d_variants_c4.txt
