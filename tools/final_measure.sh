#!/bin/bash
# Measurement pass on one B200 (run through gpurun): tests, smoke, bench (both arms), ncu launch list + full captures of
# the dominant kernels, comparison with the reference's own binaries. Output under gpurun_out/ (copy into profiles/).
R=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_$R.err | tail -1 > gpurun_out/bench_$R.json; cut -c1-400 gpurun_out/bench_$R.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference_$R.json; cut -c1-300 gpurun_out/bench_reference_$R.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4 --no-ref-gpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_s16_kernel -s 3 -c 1 -f -o gpurun_out/prof_s16_$R python tools/ncu_target.py 1000000 256 12 1 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_s16_kernel -s 3 -c 1 -f -o gpurun_out/prof_s16_multi_$R python tools/ncu_target.py 200000 768 12 1 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_s16_long2_kernel -s 0 -c 1 -f -o gpurun_out/prof_s16_long_$R python tools/ncu_long_target.py 2 2>&1 | tail -2
timeout 900 python tools/compare_reference.py c2 c2d c3 c5 2>&1 | tail -3
cp gpurun_out/compare_reference.md gpurun_out/compare_reference_$R.md
head -12 gpurun_out/compare_reference_$R.md
