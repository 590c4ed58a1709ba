mkdir -p gpurun_out; rm -f gpurun_out/r2k_pb.txt
for B in 32 64 96 130 150 512; do for b in pb_cw128_it4 pb_cw32_it4 pb_cw128 pb_cw32; do echo "== $b B=$B" >> gpurun_out/r2k_pb.txt; timeout 120 tools/bin/$b 4096 $B 2>&1 | grep -v exp_neg >> gpurun_out/r2k_pb.txt; done; done
grep -E "==|best" gpurun_out/r2k_pb.txt | cut -c1-140
timeout 600 python -m pytest tests -m gpu -q -x -k "kernel_switch or batched_rollouts_vs or randomized or dimension_sweep_both or determinism or smoke or tiny or config5" > gpurun_out/r2k_pytest.txt 2>&1; tail -3 gpurun_out/r2k_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/b1_eval.py 8 4096 30 > gpurun_out/r2k_b1.txt 2>&1; cat gpurun_out/r2k_b1.txt
timeout 600 python bench.py --config 5 --instances 2048 > gpurun_out/r2k_cfg5_n1.txt 2>&1; tail -1 gpurun_out/r2k_cfg5_n1.txt | cut -c1-300
bash tools/runs/r2j.sh
