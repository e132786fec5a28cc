mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"partition_kernel|dedup_scan" -s 2 -c 2 -o gpurun_out/prof_pd -f python scripts/prof_count_all.py 2e7 > gpurun_out/g4.log 2>&1
tail -2 gpurun_out/g4.log
