# gpurun --gpus 8 --timeout 1200 -- 'bash tools/r2_multi.sh 8'   (BASELINE configs 3 and 5 on N GPUs of one box)
set -x
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $RUN --master-port 29513 bench.py --gpus $N --config 3 --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r2z_c3_n$N.err | tail -1 > gpurun_out/r2z_bench_c3_n$N.json; cut -c1-1200 gpurun_out/r2z_bench_c3_n$N.json; tail -3 gpurun_out/r2z_c3_n$N.err
timeout 500 $RUN --master-port 29515 bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r2z_c5_n$N.err | tail -1 > gpurun_out/r2z_bench_c5_n$N.json; cut -c1-1200 gpurun_out/r2z_bench_c5_n$N.json; tail -3 gpurun_out/r2z_c5_n$N.err
