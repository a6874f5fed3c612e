set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
U2PL_CHAIN_TIMING=1 run 60 "chain dbg1" python tools/chain_time.py > $OUT/r2t_chain_time_dbg1.txt 2>&1
grep "entropy_chain. " $OUT/r2t_chain_time_dbg1.txt | tail -3
for pz in 1 2 4; do
U2PL_INFONCE_PERSIST=$pz run 100 "contra_bench persist $pz" python tools/contra_bench.py > $OUT/r2t_contra_bench_persist$pz.txt 2>&1
grep "infonce_fwd" $OUT/r2t_contra_bench_persist$pz.txt | head -1
done
U2PL_INFONCE_PERSIST=2 run 200 "pytest contra persist" python -m pytest tests/test_gpu_contra.py -q > $OUT/r2t_pytest_contra_persist.log 2>&1
tail -2 $OUT/r2t_pytest_contra_persist.log
