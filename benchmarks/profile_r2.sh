#!/bin/bash
# ncu captures of round 2 (run on the GPU box: gpurun -- bash benchmarks/profile_r2.sh); outputs in gpurun_out/
set -u
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > $O/prof_bench.json 2> $O/prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > $O/ncu_launches.log 2>&1
$NCU -k regex:edge_check_fast_kernel -s 4 -c 1 -o $O/r2_fast python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > $O/ncu_fast.log 2>&1
$NCU -k regex:edge_check_fast_kernel -s 1 -c 1 -o $O/r2_indexed python benchmarks/profile_kernels.py indexed > $O/ncu_indexed.log 2>&1
$NCU -k regex:edge_check_image -s 1 -c 1 -o $O/r2_image python benchmarks/profile_kernels.py image_indexed > $O/ncu_image.log 2>&1
$NCU -k regex:bk_pfree_fused -s 1 -c 1 -o $O/r2_pfree_fused python benchmarks/profile_kernels.py pfree_fused > $O/ncu_pfree.log 2>&1
$NCU -k regex:bk_knn_partial -s 1 -c 1 -o $O/r2_belief python benchmarks/profile_kernels.py belief > $O/ncu_belief.log 2>&1
ls -la $O
