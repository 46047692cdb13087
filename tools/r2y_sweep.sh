#!/bin/bash
# One GPU call: previous build ("base") against the current one, per kernel, on configs 2 and 4; which file a difference in
# the output bytes comes from (only when there is one); GPU tests and the bench line of the current build.
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tests/variant_sweep.py c2 base default 2>&1 | grep -v Warning | tee $O/r2y_sweep_c2.txt
if [ "$(grep -o 'crc [0-9a-f]*' $O/r2y_sweep_c2.txt | sort -u | wc -l)" != "1" ]; then
  timeout 600 python tests/variant_sweep.py c2 seq lit walk merge 2>&1 | grep -v Warning | tee -a $O/r2y_sweep_c2.txt
fi
timeout 300 python tests/variant_sweep.py c2 lit1 lit256 seq1k walk8 2>&1 | grep -v Warning | tee -a $O/r2y_sweep_c2.txt
timeout 600 python tests/variant_sweep.py c4 base default 2>&1 | grep -v Warning | tee $O/r2y_sweep_c4.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r2y_gputests.txt
timeout 600 python bench.py > $O/r2y_bench.json 2> $O/r2y_bench.err; tail -3 $O/r2y_bench.err; cat $O/r2y_bench.json
