# launch list of one evaluation at B=128 (the per-GPU shard of the 8-GPU headline run)
mkdir -p gpurun_out
CMD="python bench.py --B 128 --steps 1 --warmup 3 --no-cpu-baseline --no-latency --no-extras"
$CMD > gpurun_out/b128_plain.txt 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 140 --csv --log-file gpurun_out/b128_launches.csv $CMD > gpurun_out/b128_ncu.log 2>&1
tail -1 gpurun_out/b128_plain.txt | cut -c1-300
