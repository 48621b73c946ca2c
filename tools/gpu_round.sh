#!/bin/bash
# One GPU visit: tests, smoke, bench, ncu launch list of the bench command, ncu --set full of the top kernels.
# usage: tools/gpu_round.sh <tag> [stages...]   stages: test smoke bench launches full
mkdir -p gpurun_out
TAG=${1:-x}; shift
STAGES=${@:-test smoke bench launches full}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.log 2>&1
for s in $STAGES; do
case $s in
test)
  timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" ;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log ;;
bench)
  timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_n1.json ;;
benchref)
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "benchref rc=$?"; cat gpurun_out/${TAG}_bench_ref.json ;;
launches)
  timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_plain_bench.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "launches rc=$?" ;;
full)
  # the heavy kernels of ONE timed bench step (two halves x sweep, refine+filter, RANSAC; 3 warm-up steps x 6 launches are skipped)
  timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_plain_bench2.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'match_tc_kernel|ransac_f_kernel|refine_kernel' -s 18 -c 6 \
      -o gpurun_out/${TAG}_prof_bench -f python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_bench2.log 2>&1; echo "full rc=$?" ;;
scale8)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 \
      > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err; echo "scale8 rc=$?"; cat gpurun_out/${TAG}_bench_n8.json; tail -3 gpurun_out/${TAG}_bench_n8.err ;;
scale2)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
      > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "scale2 rc=$?"; cat gpurun_out/${TAG}_bench_n2.json; tail -3 gpurun_out/${TAG}_bench_n2.err ;;
esac
done
