# ncu counters of one launch of every kernel of a cycle, per launch size (feeds profiles/ncu_kernel_counters.json via
# scripts/counters_to_json.py): warp instructions, DRAM bytes, duration.  Run under gpurun (one GPU).
O=gpurun_out/counters; mkdir -p $O
M=smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size,launch__block_size
run() { # name, launch-skip, count, args...
  n=$1; skip=$2; cnt=$3; shift 3
  ncu --metrics $M --clock-control none --launch-skip $skip -c $cnt --csv --log-file $O/$n.csv python scripts/run_workload.py "$@" > $O/$n.log 2>&1
}
# 30 warm cycles first so that the path critics are active (steady state), then one cycle is captured (stream layout on one
# rank: three kernels per cycle - rollout, path costs, weighted sums with the merge inside)
run omni_1000x56 30 1 --workload omni_1000x56 --cycles 32 --resident
run obstacles_16384x56 90 3 --workload obstacles_16384x56 --cycles 32 --resident
run omni_262144x100 90 3 --workload sharded_262144x100 --cycles 32 --resident
run omni_131072x100 90 3 --workload sharded_262144x100 --batch 131072 --cycles 32 --resident
run omni_65536x100 90 3 --workload sharded_262144x100 --batch 65536 --cycles 32 --resident
run omni_32768x100 90 3 --workload sharded_262144x100 --batch 32768 --cycles 32 --resident
tail -n 2 $O/*.log
