mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "two_devices or split_over_two" > gpurun_out/r2u_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.txt
tail -4 gpurun_out/r2u_pytest.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --config 7 --gpus 2 > gpurun_out/r2u_cfg7_n2.txt 2>&1
tail -n 2 gpurun_out/r2u_cfg7_n2.txt | cut -c1-900
