#!/bin/bash
# tools/round2_validate.sh -- ONE gpurun call that validates everything written after round 1's GPU minutes ran out
# and collects the numbers into gpurun_out/ (copy what matters into profiles/ afterwards).
#   1 GPU :  gpurun --timeout 1500 -- 'bash tools/round2_validate.sh 1'
#   2 GPUs:  gpurun --gpus 2 --timeout 900 -- 'bash tools/round2_validate.sh 2'
set -u
mkdir -p gpurun_out
OUT=gpurun_out
if [ "${1:-1}" = "1" ]; then
  # torch-free, seconds: correctness + CUDA-event timings of the tensor-core kernels (run tools/build_selftest.sh first, here)
  if [ -x tools/cu/tc_selftest.bin ]; then
    timeout 120 ./tools/cu/tc_selftest.bin all  > $OUT/r2_tc_selftest.txt 2>&1; echo "tc_selftest: $?"
    timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2_tc_selftest_perf.txt 2>&1; echo "tc_selftest perf: $?"
    timeout 20 ./tools/cu/umma_2cta_probe.bin   > $OUT/r2_umma_2cta_probe.txt 2>&1; echo "2-CTA probe: $? (124 = hang: protocol error)"
  fi
  timeout 600 python -m pytest tests -m gpu -x -q                                   > $OUT/r2_pytest_gpu.log 2>&1;      echo "pytest default: $?"
  U2PL_TC_CONV=1 timeout 300 python -m pytest tests/test_gpu_conv_tc.py -q          > $OUT/r2_pytest_conv_tc.log 2>&1;  echo "pytest conv_tc: $?"
  timeout 200 python tools/conv_bench.py                                            > $OUT/r2_conv_bench.jsonl 2>$OUT/r2_conv_bench.err; echo "conv_bench: $?"
  timeout 300 python bench.py --impl eager --steps 5 --warmup 3                     > $OUT/r2_bench_eager.json 2>$OUT/r2_bench_eager.err; echo "bench eager: $?"
  timeout 300 python bench.py --steps 10 --warmup 3 --phases                        > $OUT/r2_bench_n1.json 2>$OUT/r2_bench_n1.err;       echo "bench ours: $?"
  U2PL_TC_CONV=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2_bench_n1_tcconv.json 2>$OUT/r2_bench_n1_tcconv.err; echo "bench ours+tc_conv: $?"
  U2PL_TC_CONV=1 U2PL_TC_TRAIN=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2_bench_n1_tctrain.json 2>$OUT/r2_bench_n1_tctrain.err; echo "bench ours+tc_conv+tc_train: $?"
  timeout 200 bash tools/umma_probe_sweep.sh > $OUT/r2_umma_probe.txt 2>&1; echo "umma MN-major probe: $? (0 = expected descriptor encoding confirmed)"
  U2PL_TC_CONV=1 U2PL_TC_WGRAD=1 timeout 300 python -m pytest tests/test_gpu_conv_tc.py -q -k wgrad > $OUT/r2_pytest_wgrad_tc.log 2>&1; echo "pytest wgrad_tc: $?"
  for d in 1 2 4; do
    U2PL_INFONCE_DEPTH=$d timeout 120 python tools/contra_bench.py > $OUT/r2_contra_bench_depth$d.log 2>&1; echo "contra_bench depth $d: $?"
  done
  U2PL_INFONCE_DEPTH=4 timeout 200 python -m pytest tests/test_gpu_contra.py -q > $OUT/r2_pytest_contra_depth4.log 2>&1; echo "pytest contra depth4: $?"
  U2PL_TC_CONV=1 U2PL_TC_CHAIN=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2_bench_n1_tcchain.json 2>$OUT/r2_bench_n1_tcchain.err; echo "bench ours+tc_conv+tc_chain: $?"
  U2PL_TC_CONV=1 U2PL_TC_TRAIN=1 U2PL_TC_CHAIN=1 U2PL_TC_WGRAD=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2_bench_n1_tcall.json 2>$OUT/r2_bench_n1_tcall.err; echo "bench ours, every tensor-core path: $?"
  U2PL_POOL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2_bench_n1_pool.json 2>$OUT/r2_bench_n1_pool.err; echo "bench ours+pool: $?"
  U2PL_WGRAD_STACK=1 timeout 200 python -m pytest tests/test_gpu_fused.py -q -k dilated > $OUT/r2_pytest_wgrad_stack.log 2>&1; echo "pytest wgrad_stack: $?"
  U2PL_WGRAD_STACK=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2_bench_n1_wgradstack.json 2>$OUT/r2_bench_n1_wgradstack.err; echo "bench ours+wgrad_stack: $?"
else
  N=$1
  U2PL_BANK_SHARDED_TEST=1 timeout 300 python -m pytest tests/test_gpu_sharded_bank.py -x -q > $OUT/r2_pytest_sharded.log 2>&1; echo "pytest sharded: $?"
  for mode in 0 1; do
    U2PL_BANK_SHARDED=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2_bench_n${N}_sharded${mode}.json 2>$OUT/r2_bench_n${N}_sharded${mode}.err
    echo "bench N=$N sharded=$mode: $?"
  done
fi
tail -n 3 $OUT/r2_*.log 2>/dev/null
# profiling (run separately, after the tests above are green):
#   gpurun --timeout 600 -- 'ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 1 -o gpurun_out/r02_conv_tc python tools/conv_one.py layer3'
#   gpurun --timeout 600 -- 'U2PL_TC_CONV=1 U2PL_TC_TRAIN=1 python tools/step_profile.py > gpurun_out/r02_step_profile.txt'
#   full ncu launch list of one bench step (≈ 20 GPU-min for ~6000 launches; round 1's attempt was cut at 2517 launches):
#   gpurun --timeout 1500 -- 'U2PL_BENCH_FAST=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_step_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1'
