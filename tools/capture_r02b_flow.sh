set -u
NCU="ncu --set full --clock-control none --import-source on -f"
B="python bench.py --no-nested --no-cpu-baseline --steps 1 --warmup 3"
$B > gpurun_out/r02b_pre_ncu.json 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02b_launches_lazy.csv $B > gpurun_out/cap_launches.log 2>&1
timeout 600 $NCU -k regex:predict_flow_kernel -s 6 -c 1 -o gpurun_out/flow_cluster_r02b $B > gpurun_out/cap_flowc.log 2>&1
tail -n 2 gpurun_out/cap_launches.log | cut -c1-200; tail -n 2 gpurun_out/cap_flowc.log | cut -c1-200
python __graft_entry__.py smoke 2>&1 | tail -n 1
