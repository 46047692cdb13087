timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_r1_n2.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r1_n2.json")); print({k:d[k] for k in ("value","n_gpus","ms_per_step","e2e","gpu_launches","clocks")})
PY
timeout 300 python -m pytest tests -m gpu -x -q -k "many or dict or frames" 2>&1 | tail -2
