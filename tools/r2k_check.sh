#!/bin/bash
# round-2 check after the splitter fix: GPU test suite, bench line, imbalance of tiled crops, 1,000-replay soak at 65,536 rays
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2k_pytest.log
timeout 300 python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo "bench rc=$?"; head -c 400 gpurun_out/r2k_bench_n1.json; echo
for cr in 5 6; do
IRONB_BENCH_CROP_RANK=$cr IRONB_BENCH_CROP_STRIDE=64 timeout 120 python bench.py --no-cpu --no-clocks --steps 10 > gpurun_out/r2k_tiled_crop$cr.json 2> gpurun_out/r2k_tiled_crop$cr.err; echo "tiled crop $cr rc=$?"; head -c 200 gpurun_out/r2k_tiled_crop$cr.json; echo
IRONB_BENCH_CROP_RANK=$cr timeout 120 python bench.py --no-cpu --no-clocks --steps 10 > gpurun_out/r2k_near_crop$cr.json 2> gpurun_out/r2k_near_crop$cr.err; echo "near crop $cr rc=$?"; head -c 200 gpurun_out/r2k_near_crop$cr.json; echo
done
REPRO_STEPS=1000 REPRO_TIMEOUT=200 IRONB_BENCH_WATCHDOG_S=190 tools/repro_p256.sh r2k_soak1000
