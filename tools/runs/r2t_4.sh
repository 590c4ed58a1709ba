mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2t_bench_n4.txt 2>&1
timeout 600 $TR bench.py --config 5 --gpus 4 --instances 2048 > gpurun_out/r2t_cfg5_n4.txt 2>&1
tail -n 1 gpurun_out/r2t_bench_n4.txt gpurun_out/r2t_cfg5_n4.txt | cut -c1-400
