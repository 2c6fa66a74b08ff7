set -x
O=gpurun_out/r01c; mkdir -p $O
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv -lms 100 > $O/clocks_during_bench_default.csv 2>/dev/null &
SMI=$!
python bench.py > $O/bench_omni_1000x56.json 2> $O/bench_omni_1000x56.err
kill $SMI
python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --workload obstacles_16384x56 > $O/bench_obstacles_16384x56.json 2> $O/bench_obstacles.err
python bench.py --workload sharded_262144x100 --steps 200 > $O/bench_sharded_262144x100_1gpu.json 2> $O/bench_sharded.err
python bench.py --workload robots_256 --steps 50 > $O/bench_robots_256_1gpu.json 2> $O/bench_robots.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_default.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
python scripts/run_workload.py --workload omni_1000x56 --cycles 8 --resident > $O/plain_small.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tile_fused -s 5 -c 1 -o $O/prof_fused -f python scripts/run_workload.py --workload omni_1000x56 --cycles 8 --resident > $O/ncu_small.log 2>&1
ls -la $O
