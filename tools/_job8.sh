cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2b_bench_n8.json 2> gpurun_out/r2b_bench_n8.err
tail -c 3000 gpurun_out/r2b_bench_n8.json
