#!/bin/bash
# One GPU-box session: parity tests, kernel-variant microbench, bench, config sweep, ncu launch list + full capture.
# Usage: bash tools/gpu_session.sh [steps...]   (default: all)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
STEPS="${*:-smoke pytest kbench bench configs ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
for s in $STEPS; do
case $s in
smoke)
  python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log ;;
pytest)
  timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log ;;
kbench)
  KBENCH_VARIANTS="${KBENCH_VARIANTS:-0,2,7,9,14,15,16,17,18,19,20}" timeout 600 python tools/kbench.py > gpurun_out/kbench.log 2>&1 ;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 3 --profile-host > gpurun_out/bench.log 2> gpurun_out/bench_hostprofile.log; echo "bench rc=$?" >> gpurun_out/bench.log ;;
benchab)
  for b in 1 0; do
    FLOXB_WORKSPACE=$b timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_ws$b.log 2>&1
  done ;;
ncukexp)
  KBENCH_VARIANTS=22 KBENCH_GROUPS=10000 timeout 600 ncu --set full --clock-control none -k regex:k_exp_pipe -s 1 -c 1 -f -o gpurun_out/prof_kexp22 \
      python tools/kbench.py 268435456 > gpurun_out/ncu_kexp.log 2>&1 ;;
kbench3)
  timeout 600 python tools/kbench3.py > gpurun_out/kbench3.log 2>&1 ;;
ncusecond)
  for pair in rows:k_rows_bins c5:k_moments_global_runs c3:k_cols; do
    cfg=${pair%%:*}; kern=${pair##*:}
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -c 2 -f -o gpurun_out/prof_$kern \
        python tools/configs_bench.py $cfg > gpurun_out/ncu_$kern.log 2>&1
  done ;;
ncuother)
  timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:k_moments_global|k_rows|k_sqdev_global' -c 8 -f -o gpurun_out/prof_other \
      python tools/configs_bench.py c4 c5 rows > gpurun_out/ncu_other.log 2>&1 ;;
e2eprobe)
  timeout 600 python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1 ;;
kbench2)
  timeout 600 python tools/kbench2.py > gpurun_out/kbench2.log 2>&1 ;;
configs)
  timeout 900 python tools/configs_bench.py > gpurun_out/configs.log 2>&1 ;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_moments_smem -s 1 -c 2 -f -o gpurun_out/prof_moments \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1 ;;
ncuconfigs)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_configs.csv \
      -k regex:^k_ python tools/configs_bench.py c4 c5 rows > gpurun_out/ncu_configs.log 2>&1 ;;
ncucols)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cols -s 1 -c 1 -f -o gpurun_out/prof_cols \
      python tools/configs_bench.py c3 > gpurun_out/ncu_cols.log 2>&1 ;;
r2final)  # round-2 final single-GPU session: smoke, the whole -m gpu tier, the driver-shaped bench, the reference arm, ncu
  python __graft_entry__.py --smoke > gpurun_out/s10_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s10_smoke.log
  timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/s10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s10_pytest.log
  timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/s10_bench.log 2>&1; echo "rc=$?" >> gpurun_out/s10_bench.log
  timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/s10_bench_ref.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s10_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s10_ncu_launches.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:bruns -s 2 -c 2 -f -o gpurun_out/prof_c4 \
      python tools/configs_bench.py c4 > gpurun_out/s10_ncu_c4.log 2>&1 ;;
ncufx)    # the headline kernel: bench line + ncu --set full
  timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/s3_bench_fx.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_moments_smem_fx -s 1 -c 1 -f -o gpurun_out/prof_fx \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s3_ncu.log 2>&1 ;;
c4ab)     # config 4 with the run-merging kernels and with the round-1 kernels (FLOXB_FORCE_PATH=4)
  timeout 600 python tools/configs_bench.py c4 > gpurun_out/s2_c4_new.log 2>&1
  FLOXB_FORCE_PATH=4 timeout 600 python tools/configs_bench.py c4 > gpurun_out/s2_c4_old.log 2>&1 ;;
breakdown)
  timeout 600 python tools/step_breakdown.py 1000000000 > gpurun_out/s5_bd1.log 2>&1
  timeout 600 python tools/step_breakdown.py 125000000 > gpurun_out/s5_bd2.log 2>&1 ;;
multi)    # run with gpurun --gpus N: parity worker, bench (strong + weak + e2e), configs 3/4/5 on N GPUs
  NG=$(nvidia-smi -L | wc -l)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
      tests/dist_worker.py > gpurun_out/s9_dist_$NG.log 2>&1; echo "rc=$?" >> gpurun_out/s9_dist_$NG.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/s9_bench_$NG.log 2>&1; echo "rc=$?" >> gpurun_out/s9_bench_$NG.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 \
      tools/configs_multi.py ${S9_CONFIGS:-c3 c4 c5} > gpurun_out/s9_configs_$NG.log 2>&1 ;;
dist)
  NG=$(nvidia-smi -L | wc -l)
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
      tests/dist_worker.py > gpurun_out/dist_check.log 2>&1; echo "dist_worker rc=$?" >> gpurun_out/dist_check.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/bench_multi.log 2>&1; echo "bench_multi rc=$?" >> gpurun_out/bench_multi.log ;;
esac
done
ls -la gpurun_out
tail -4 gpurun_out/pytest_gpu.log gpurun_out/bench.log gpurun_out/smoke.log gpurun_out/configs.log 2>/dev/null
