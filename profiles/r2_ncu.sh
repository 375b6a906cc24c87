#!/bin/bash
# Round 2: ncu launch list + one --set full capture of two steady-state frames (frames 7 and 8 of 8) of the default build.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_launches_r2.csv python profiles/run_profile.py 8 > gpurun_out/ncu_list_r2.log 2>&1; tail -1 gpurun_out/ncu_list_r2.log
ncu --set full --clock-control none --import-source on -k regex:"tsdfIntegrate|esdfMarkTma|esdfClear|esdfWaveX|viewRaycast|compactAllocate|esdfAllocate" -s 42 -c 14 -o gpurun_out/prof_r2 -f python profiles/run_profile.py 8 > gpurun_out/ncu_full_r2.log 2>&1; tail -2 gpurun_out/ncu_full_r2.log
ncu -i gpurun_out/prof_r2.ncu-rep --page raw --csv > gpurun_out/ncu_full_raw_r2.csv 2>/dev/null
ls -la gpurun_out/prof_r2.ncu-rep gpurun_out/ncu_full_raw_r2.csv
