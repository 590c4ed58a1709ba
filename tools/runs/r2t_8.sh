mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2t_bench_n8.txt 2>&1
timeout 600 $TR bench.py --config 5 --gpus 8 --instances 2048 > gpurun_out/r2t_cfg5_n8.txt 2>&1
tail -n 1 gpurun_out/r2t_bench_n8.txt gpurun_out/r2t_cfg5_n8.txt | cut -c1-400
