#!/bin/bash
# Runs on the GPU box: the measurement set behind profiles/ (bench both arms, per-op sweep vs the reference, ncu launch list of the
# bench command, one ncu --set full capture per kernel family). Outputs go to gpurun_out/$1_*.
tag=${1:-rX}
mkdir -p gpurun_out
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py > gpurun_out/${tag}_bench_ours.json 2> gpurun_out/${tag}_bench_ours.err
python scripts/bench_ops.py --ops cholsweep,gemm,gemmsweep,gels,qr,svd,svdu,nullspace --ref --json gpurun_out/${tag}_ops.json > gpurun_out/${tag}_ops.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_potrf_group|k_potrs_group" -s 2 -c 2 -o gpurun_out/${tag}_prof_chol32 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --batch 250000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_geqrf_tc|k_gels_f2" -c 2 -o gpurun_out/${tag}_prof_qr python scripts/bench_ops.py --ops qr,gels --reps 1 --scale 0.25 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_jacobi_rt|k_ormqr_tc|k_gemv|k_gemm_dmma" -c 6 -o gpurun_out/${tag}_prof_svd python scripts/bench_ops.py --ops svdu,nullspace --reps 1 --scale 0.125 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gemm_col|k_sgemm32_rt|k_dgemm_frag" -s 3 -c 1 -o gpurun_out/${tag}_prof_gemm8 python scripts/bench_ops.py --ops gemm --reps 1 > /dev/null 2>&1
