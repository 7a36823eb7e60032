#!/bin/bash
# Final evidence of the round: launch list of one bench step, DRAM traffic and full captures of K-fwd / K-inv
# on the bench's own batch (4096 images on one GPU), entropy kernels at 512 images, SASS summary.
tag=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-subrecords"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_(forward|inverse|huff|dec|lowres|lres|inv_tables|store)" -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:k_forward3|k_inverse4" -c 6 --csv --log-file gpurun_out/${tag}_traffic.csv $B > gpurun_out/${tag}_traffic.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:k_forward3|k_inverse4" -c 2 -o gpurun_out/${tag}_xform -f python tools/kernel_times.py 512 50 > gpurun_out/${tag}_xform.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:k_huff_pack3|k_huff_hist2|k_dec_stream_par|k_lowres_avg" -c 5 -o gpurun_out/${tag}_entropy -f python tools/kernel_times.py 512 50 > gpurun_out/${tag}_entropy.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:k_dec_stream_par" -c 2 -o gpurun_out/${tag}_dec -f python tools/kernel_times.py 512 50 > gpurun_out/${tag}_dec.log 2>&1
ls -la gpurun_out | tail -12
