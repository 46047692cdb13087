set -x
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $RUN --master-port 29515 bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r2z_c5_n$N.err | grep '^{' | tail -1 > gpurun_out/r2z_bench_c5_n$N.json; cut -c1-1200 gpurun_out/r2z_bench_c5_n$N.json; tail -n 3 gpurun_out/r2z_c5_n$N.err
