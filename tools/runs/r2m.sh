mkdir -p gpurun_out
tools/bin/pb2_cw32 4096 64 > gpurun_out/r2m_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mm_pairs_batch -s 1 -c 1 -o gpurun_out/r2m_cw32 tools/bin/pb2_cw32 4096 64 > gpurun_out/r2m_ncu.log 2>&1
tail -2 gpurun_out/r2m_ncu.log
