mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"partition|bucket|dedup|terminal" --csv --log-file gpurun_out/l3.csv python scripts/prof_count_all.py 1e8 > gpurun_out/l3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"partition|bucket|dedup|terminal" --csv --log-file gpurun_out/l3rep.csv python scripts/prof_count_all.py 1e8 rep > gpurun_out/l3rep.log 2>&1
python scripts/summarise_launches.py gpurun_out/l3.csv; python scripts/summarise_launches.py gpurun_out/l3rep.csv
