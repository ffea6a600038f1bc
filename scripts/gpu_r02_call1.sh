#!/bin/bash
# round 2, GPU call 1: baseline of the round-1 kernels on the stress configs + eigenvalue histogram + CUDASIM probe
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_c1_tests.log
python scripts/eigen_error_hist.py --impl cuda --out gpurun_out/eigen_hist_r02_cuda.json > gpurun_out/r02_c1_eig.log 2>&1
python bench.py --steps 100 --warmup 20 > gpurun_out/bench_r02_base_n1.json 2> gpurun_out/bench_r02_base_n1.err
python bench.py --config long_range --steps 60 --warmup 20 > gpurun_out/bench_r02_base_long_range.json 2> gpurun_out/bench_r02_base_long_range.err
python bench.py --config dense --steps 12 --warmup 4 > gpurun_out/bench_r02_base_dense.json 2> gpurun_out/bench_r02_base_dense.err
(time NUMBA_ENABLE_CUDASIM=1 python baseline/ref_probe.py --xy 16 --z 8 --beams 4 --cols 24 --iters 1 --out gpurun_out/ref_probe_cudasim_box.json) > gpurun_out/r02_c1_cudasim.log 2>&1
nproc >> gpurun_out/r02_c1_cudasim.log
tail -3 gpurun_out/r02_c1_tests.log
