mkdir -p gpurun_out; rm -f gpurun_out/r2p2_pb.txt
for b in pb_rs22 pb_xs pb_rs22 pb_xs; do echo "== $b" >> gpurun_out/r2p2_pb.txt; timeout 120 tools/bin/$b 4096 1024 2>&1 | grep -v "exp_neg" >> gpurun_out/r2p2_pb.txt; done
cat gpurun_out/r2p2_pb.txt
