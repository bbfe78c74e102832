#!/bin/bash
# Round-end measurement pass on one B200 (run through gpurun): tests, bench, ncu launch list + full captures, comparison.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r1.json; cat gpurun_out/bench_r1.json | cut -c1-300
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference_r1.json; cat gpurun_out/bench_reference_r1.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_s16_kernel -s 2 -c 1 -f -o gpurun_out/prof_s16_r1_final python tools/ncu_target.py 1000000 256 9 1 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_s16_long_kernel -s 2 -c 1 -f -o gpurun_out/prof_s16_long_r1 python tools/ncu_long_target.py 2 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_s32_long_kernel -s 1 -c 1 -f -o gpurun_out/prof_s32_long_r1 python tools/ncu_long_target.py 2 2>&1 | tail -2
timeout 900 python tools/compare_reference.py c2 c2d c3 c5 2>&1 | tail -3
head -12 gpurun_out/compare_reference.md
