#!/bin/bash
# Runs on the GPU box: the measurement set behind profiles/ (bench both arms, per-op sweep vs the reference, ncu launch list of the
# bench command, one ncu --set full capture per kernel family). Outputs go to gpurun_out/$1_*.
tag=${1:-rX}
mkdir -p gpurun_out
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py > gpurun_out/${tag}_bench_ours.json 2> gpurun_out/${tag}_bench_ours.err
python scripts/bench_ops.py --ops cholsweep,gemm,gemmsweep,gels,qr,svd,svdu,nullspace --ref --json gpurun_out/${tag}_ops.json > gpurun_out/${tag}_ops.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:"k_potrf_pair|k_potrs_group|k_potrs_pair" -s 2 -c 2 -o gpurun_out/${tag}_prof_chol32 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --batch 250000 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_geqrf_tc" -c 1 -o gpurun_out/${tag}_prof_geqrf python scripts/bench_ops.py --ops qr --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_gels_f2" -c 1 -o gpurun_out/${tag}_prof_gels python scripts/bench_ops.py --ops gels --reps 1 --scale 0.25 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_jacobi_rt|k_jacobi_blk|k_ormqr_tc|k_gemv|k_gemm_dmma" -c 8 -o gpurun_out/${tag}_prof_svd python scripts/bench_ops.py --ops svdu,nullspace --reps 1 --scale 0.125 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_gemm_col" -s 5 -c 1 -o gpurun_out/${tag}_prof_gemm8 python scripts/bench_ops.py --ops gemm --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_potrf_pair|k_potrs_pair64|k_potrf_pipe|k_potrs_quad128" -c 8 -o gpurun_out/${tag}_prof_chol64 python scripts/bench_ops.py --ops cholsweep --reps 1 --scale 0.125 > /dev/null 2>&1
# keep what travels back small (gpurun merges at most 64 MiB): summarise every capture here and drop the .ncu-rep
for f in gpurun_out/${tag}_prof_*.ncu-rep; do
    python scripts/ncu_raw.py $f --json ${f%.ncu-rep}.json > /dev/null 2>&1
    rm -f $f
done
