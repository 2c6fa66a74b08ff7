# weak-scaling line of the default workload at N GPUs: bash scripts/collect_weak.sh N   (under gpurun --gpus N)
N=$1
O=gpurun_out/r01c; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > $O/bench_omni_1000x56_${N}gpu_weak.json 2> $O/err_weak_$N.log
tail -n 3 $O/err_weak_$N.log | cut -c1-300
cut -c1-300 $O/bench_omni_1000x56_${N}gpu_weak.json
