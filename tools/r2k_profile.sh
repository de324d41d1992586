#!/bin/bash
# round-2 final profiles: launch list of the bench step, ncu --set full of the three tcgen05 kernels (each after its plain run)
mkdir -p gpurun_out
tools/launch_list.sh r2k
NCU="ncu --set full --clock-control none --import-source on"
python tests/profile_trace.py --mode 2 --reps 1 > gpurun_out/plain_prof_trace.log 2>&1 && \
$NCU -k regex:mlp_h16 -s 0 -c 2 -f -o gpurun_out/prof_mlp_r2k python tests/profile_trace.py --mode 2 --reps 1 > gpurun_out/ncu_prof_mlp_r2k.log 2>&1
echo "mlp_h16 full rc=$?"
export IRONB_BENCH_MIN_WARMUP_S=0
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-clocks"
$NCU -k regex:gemm_h16_kernel -s 200 -c 4 -f -o gpurun_out/prof_gemmh16_r2k $B > gpurun_out/ncu_prof_gemmh16_r2k.log 2>&1
echo "gemm_h16 full rc=$?"
$NCU -k regex:gemm_nt_tc_kernel -s 300 -c 6 -f -o gpurun_out/prof_gemmtc_r2k $B > gpurun_out/ncu_prof_gemmtc_r2k.log 2>&1
echo "gemm_nt_tc full rc=$?"
ls -la gpurun_out/*.ncu-rep
