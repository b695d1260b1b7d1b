#!/bin/bash
# ncu captures of round 2 (run on the GPU box: gpurun -- 'bash tools/capture_r02.sh'); reports land in gpurun_out/
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
B="python bench.py --no-nested --no-cpu-baseline --steps 1 --warmup 3"
# dataflow forecast kernel (lazy workload: 16 groups x N=2000); the 3rd launch of the run = steps 513..768 of a forecast
timeout 400 $NCU -k regex:predict_flow_kernel -s 6 -c 1 -o gpurun_out/flow_r02_final $B > gpurun_out/cap_flow.log 2>&1
# eager N=16000: SYRK of the ridge solve on the int8 tensor cores, the diagonal-block factorisation, the barrier-free forecast
E="python bench.py --workload eager16k --no-cpu-baseline --steps 1 --warmup 3"
timeout 500 $NCU -k regex:gram_i8_kernel -s 30 -c 1 -o gpurun_out/solve_syrk_i8_r02 $E > gpurun_out/cap_syrk.log 2>&1
timeout 500 $NCU -k regex:potrf_diag_kernel -s 130 -c 1 -o gpurun_out/potrf_diag_r02 $E > gpurun_out/cap_potrf.log 2>&1
timeout 500 $NCU -k regex:predict_pull_kernel -s 5 -c 1 -o gpurun_out/predict_pull_r02 $E > gpurun_out/cap_pull.log 2>&1
timeout 500 $NCU -k regex:fused_chunk_i8 -s 14 -c 1 -o gpurun_out/fused_eager16k_r02 $E > gpurun_out/cap_fused16k.log 2>&1
# cfg 4 share of one GPU: the readout that streams Wout (49.2 MB per group and step) in the per-step closed loop
L="python bench.py --workload lazy2d --no-cpu-baseline --steps 1 --warmup 3"
timeout 600 $NCU -k regex:readout_kernel -s 400 -c 1 -o gpurun_out/readout_lazy2d_r02 $L > gpurun_out/cap_readout.log 2>&1
ls -la gpurun_out/*_r02*.ncu-rep
tail -2 gpurun_out/cap_*.log
