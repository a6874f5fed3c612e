set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
for d in 5 9 13; do
U2PL_CHAIN_TIMING=$d run 60 "chain dbg$d" python tools/chain_time.py > $OUT/r2s_chain_time_dbg$d.txt 2>&1
grep "entropy_chain. grid" $OUT/r2s_chain_time_dbg$d.txt | tail -1
done
for d in 1 2 4; do
U2PL_INFONCE_DEPTH=$d run 100 "contra_bench depth $d" python tools/contra_bench.py > $OUT/r2s_contra_bench_depth$d.txt 2>&1
tail -1 $OUT/r2s_contra_bench_depth$d.txt
done
U2PL_INFONCE_DEPTH=4 run 200 "pytest contra depth4" python -m pytest tests/test_gpu_contra.py -q > $OUT/r2s_pytest_contra_depth4.log 2>&1
tail -2 $OUT/r2s_pytest_contra_depth4.log
run 400 "pytest new (-s)" python -m pytest tests/test_gpu_entropy.py::test_v16_index_set_flips_vs_torch_cuda_eager tests/test_gpu_step.py::test_v16_bf16_vs_fp32_three_losses -q -s > $OUT/r2s_pytest_flips_precision.log 2>&1
grep "^\[" $OUT/r2s_pytest_flips_precision.log
