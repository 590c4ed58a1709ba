mkdir -p gpurun_out; rm -f gpurun_out/r2l_pb.txt
for spec in "pb_cw128 1024" "pb2_cw128 1024" "pb_cw128 1024" "pb2_cw128 1024" "pb_cw128 256" "pb2_cw128 256" "pb_cw32 64" "pb2_cw32 64" "pb_cw32 130" "pb2_cw32 130" "pb_cw32 512" "pb2_cw32 512" "pb2_cw32 32" "pb2_cw32 96"; do set -- $spec; echo "== $1 B=$2" >> gpurun_out/r2l_pb.txt; timeout 120 tools/bin/$1 4096 $2 2>&1 | grep -v exp_neg >> gpurun_out/r2l_pb.txt; done
grep -E "==|best" gpurun_out/r2l_pb.txt | cut -c1-230
