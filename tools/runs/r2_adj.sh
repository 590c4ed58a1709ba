mkdir -p gpurun_out
CMD="python tools/b1_eval.py 6"
$CMD > gpurun_out/adj_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cost_adjoint_small -s 3 -c 1 -o gpurun_out/adj_small $CMD > gpurun_out/adj_ncu.log 2>&1
tail -3 gpurun_out/adj_ncu.log
