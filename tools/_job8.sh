cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2d_bench_n8.json 2> gpurun_out/r2d_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2d_bench_n4.json 2> gpurun_out/r2d_bench_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
python - <<'PY'
import json
for n in (8,4,2):
    try:
        d=json.loads(open('gpurun_out/r2d_bench_n%d.json'%n).read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d["per_rank_ms"]["min"], d["per_rank_ms"]["max"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["parity"].get("ranks_checked"), d["parity"]["max_rel"])
    except Exception as e:
        print(n, "failed", e)
PY
