# multi-GPU evidence: bash scripts/collect_evidence_multi.sh N   (N = 2, 4 or 8; run under gpurun --gpus N)
N=$1
O=gpurun_out/r01c; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
run > $O/bench_omni_1000x56_${N}gpu_weak.json 2> $O/err_weak_$N.log
run --impl reference > $O/bench_reference_arm_${N}gpu.json 2> $O/err_ref_$N.log
run --workload sharded_262144x100 --steps 200 --exchange peer > $O/bench_sharded_262144x100_${N}gpu_peer.json 2> $O/err_peer_$N.log
run --workload sharded_262144x100 --steps 200 --exchange nccl --no-cpu-baseline > $O/bench_sharded_262144x100_${N}gpu_nccl.json 2> $O/err_nccl_$N.log
run --workload robots_256 --steps 50 --no-cpu-baseline > $O/bench_robots_256_${N}gpu.json 2> $O/err_robots_$N.log
for f in $O/err_*_$N.log; do tail -n 2 $f; done
for f in $O/*_${N}gpu*.json; do echo $f; cut -c1-400 $f; done
