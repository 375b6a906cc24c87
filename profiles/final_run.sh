#!/bin/bash
# Round-end measurement set on one B200 (see profiles/README.md).
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -2 gpurun_out/bench_r1_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_launches_r1_final.csv python profiles/run_profile.py 8 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
ncu --set full --clock-control none --import-source on -k regex:"tsdfIntegrate|esdfMarkTma|esdfClear|esdfWave|viewRaycast|compactAllocate" -s 36 -c 12 -o gpurun_out/prof_r1_final python profiles/run_profile.py 8 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 300 python profiles/stress_determinism.py 3 80 async 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
