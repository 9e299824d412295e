cd /root/repo
mkdir -p gpurun_out
(nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"; python -c "import os; print(len(os.sched_getaffinity(0)))") > gpurun_out/topo.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2c_bench_n8.json 2> gpurun_out/r2c_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_n8.json').read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d.get("host_affinity"))
PY
cat gpurun_out/topo.log
