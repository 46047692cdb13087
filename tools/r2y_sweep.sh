#!/bin/bash
# One GPU call: an earlier build against the current one and its variants, per kernel (tests/variant_sweep.py), then the GPU
# tests and the bench line of the current build.   usage: tools/r2y_sweep.sh TAG "c2 variants..." "c4 variants..."
set -x
mkdir -p gpurun_out
O=gpurun_out; T=${1:-r2y}
timeout 900 python tests/variant_sweep.py c2 $2 2>&1 | grep -v Warning | tee $O/${T}_sweep_c2.txt
[ -n "$3" ] && timeout 600 python tests/variant_sweep.py c4 $3 2>&1 | grep -v Warning | tee $O/${T}_sweep_c4.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/${T}_gputests.txt
timeout 600 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; tail -3 $O/${T}_bench.err; cut -c1-900 $O/${T}_bench.json
[ -n "$4" ] && timeout 300 python tests/wave_sweep.py c2 4 2>&1 | grep -v Warning | tee $O/${T}_wave_sweep.txt
