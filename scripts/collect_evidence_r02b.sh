# round-2 (second half) evidence, one GPU: bash scripts/collect_evidence_r02b.sh   (under gpurun; outputs in gpurun_out/r02b/)
O=gpurun_out/r02b; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/test_gpu_1gpu.log
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv -lms 100 > $O/clocks_during_bench_default.csv 2>/dev/null &
SMI=$!
python bench.py > $O/bench_default_1gpu.json 2> $O/bench_default_1gpu.err
kill $SMI
python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
# launch list of the default line (the same command without ncu has exited 0 just above)
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-nested > $O/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_default.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-nested > $O/ncu_bench.log 2>&1
# full captures: the fused small-batch kernel; the three stream-layout kernels at 262144 x 100 in the steady state
python scripts/run_workload.py --workload omni_1000x56 --cycles 8 --resident > $O/plain_small.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tile_fused -s 5 -c 1 -o $O/prof_fused -f python scripts/run_workload.py --workload omni_1000x56 --cycles 8 --resident > $O/ncu_small.log 2>&1
python scripts/run_workload.py --workload sharded_262144x100 --cycles 34 --resident > $O/plain_big.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"weighted_sums_tma|rollout_score_stream|path_costs_tm" --launch-skip 90 -c 3 -o $O/prof_stream_262144x100 -f python scripts/run_workload.py --workload sharded_262144x100 --cycles 34 --resident > $O/ncu_big.log 2>&1
bash scripts/collect_counters.sh > $O/counters.log 2>&1
cat $O/test_gpu_1gpu.log; tail -n 2 $O/*.err $O/ncu_*.log
ls -la $O
