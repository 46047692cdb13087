set -x
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2z4_gpudecode.txt
timeout 300 python tests/bench_decode.py 2>&1 | tee gpurun_out/r2z4_decode_sm_scope.txt
ZSTDB200_LIB=zstd_b200/variants/libzstd_b200_ldl2.so timeout 300 python tests/bench_decode.py 2>&1 | tee gpurun_out/r2z4_decode_l2.txt
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -2
