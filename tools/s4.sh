cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/s4_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py > gpurun_out/s4_dist.log 2>&1; echo "rc=$?" >> gpurun_out/s4_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/s4_bench.log 2>&1; echo "rc=$?" >> gpurun_out/s4_bench.log
FLOXB_PEER_XCHG=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 20 --warmup 3 --no-e2e --no-other-scaling > gpurun_out/s4_bench_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/s4_bench_nccl.log
tail -5 gpurun_out/s4_dist.log; tail -2 gpurun_out/s4_bench.log | cut -c1-3500; tail -2 gpurun_out/s4_bench_nccl.log | cut -c1-1200
