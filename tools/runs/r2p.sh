mkdir -p gpurun_out
rm -f gpurun_out/r2p_pb.txt
for b in pb_ns4 pb_rs41 pb_rs22 pb_rs21 pb_rs42; do echo "== $b" >> gpurun_out/r2p_pb.txt; timeout 120 tools/bin/$b 4096 1024 2>&1 | grep -v "exp_neg" >> gpurun_out/r2p_pb.txt; done
for B in 128 256 512; do for it in 2 4 8 16 32; do echo "== items $it B $B" >> gpurun_out/r2p_pb.txt; timeout 120 tools/bin/pb_it$it 4096 $B 2>&1 | grep -v "exp_neg" >> gpurun_out/r2p_pb.txt; done; done
cat gpurun_out/r2p_pb.txt
