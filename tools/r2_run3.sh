set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2c_gputests.txt
timeout 600 python tests/variant_sweep.py c2 default pf0 pf2 pfc5 2>&1 | tee gpurun_out/r2c_variants_c2.txt
timeout 300 python tests/variant_sweep.py c4 default pf0 2>&1 | tee gpurun_out/r2c_variants_c4.txt
