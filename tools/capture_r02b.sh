#!/bin/bash
# ncu captures of the second half of round 2 (run on the GPU box: gpurun -- 'bash tools/capture_r02b.sh'); reports land
# in gpurun_out/.  Default workload (lazy: 16 groups x N=2000): the recurrence clusters and the Gram worker grid now run as
# two kernels side by side, the forecast is the cluster variant of the dataflow kernel.
set -u
mkdir -p gpurun_out
# ncu runs the two kernels of a chunk launch one after the other: keep the recurrence CTAs from taking the Gram tiles
# (in the pipeline they only share what is left when their chunk is done)
export XESN_B200_NO_JOIN=1
NCU="ncu --set full --clock-control none --import-source on -f"
B="python bench.py --no-nested --no-cpu-baseline --steps 1 --warmup 3"
# launch list first (the program has already run to completion without ncu in the calling command)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02b_launches_lazy.csv $B > gpurun_out/cap_launches.log 2>&1
# recurrence clusters (16 clusters of 4 CTAs x 500 units, 2000 steps): a launch in the middle of a training pass
timeout 600 $NCU -k regex:fused_cluster_i8_kernel -s 20 -c 1 -o gpurun_out/recur_cluster_r02b $B > gpurun_out/cap_cluster.log 2>&1
# Gram / Y R / drive workers beside them (84 CTAs)
timeout 600 $NCU -k regex:fused_chunk_i8_kernel -s 20 -c 1 -o gpurun_out/gram_workers_r02b $B > gpurun_out/cap_workers.log 2>&1
# dataflow forecast, cluster variant (16 clusters of 6)
timeout 600 $NCU -k regex:predict_flow_kernel -s 6 -c 1 -o gpurun_out/flow_cluster_r02b $B > gpurun_out/cap_flowc.log 2>&1
ls -la gpurun_out/*_r02b*.ncu-rep
for f in cluster workers flowc; do tail -n 2 gpurun_out/cap_$f.log | cut -c1-200; done
