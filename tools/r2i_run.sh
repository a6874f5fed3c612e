set -u
OUT=gpurun_out
mkdir -p $OUT
U2PL_CHAIN_TIMING=1 timeout 120 python tools/chain_time.py > $OUT/r2i_chain_time.txt 2>&1; echo "chain timing: $?"
timeout 120 python tools/chain_time.py > $OUT/r2i_chain_time_clean.txt 2>&1; echo "chain clean: $?"
C=19 timeout 120 python tools/chain_time.py > $OUT/r2i_chain_time_c19.txt 2>&1; echo "chain c19: $?"
timeout 600 python -m pytest tests/test_gpu_entropy.py tests/test_gpu_optim.py -q > $OUT/r2i_pytest_entropy.log 2>&1; echo "pytest entropy: $?"
U2PL_CONV_STAGES=3 timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_flat -s 5 -c 1 -o $OUT/r02_conv_flat_l3conv3_s3 ./tools/cu/tc_selftest.bin perf 2 > $OUT/r2i_ncu1.log 2>&1; echo "ncu conv3: $?"
U2PL_CONV_STAGES=3 timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_flat -s 18 -c 1 -o $OUT/r02_conv_flat_l3conv3_res_s3 ./tools/cu/tc_selftest.bin perf 2 > $OUT/r2i_ncu2.log 2>&1; echo "ncu conv3+res: $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_flat -s 5 -c 1 -o $OUT/r02_conv_flat_l4conv2_s4 ./tools/cu/tc_selftest.bin perf 3 > $OUT/r2i_ncu3.log 2>&1; echo "ncu l4conv2: $?"
U2PL_ENTROPY_CHAIN=0 U2PL_TC_TRAIN=1 timeout 600 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2i_bench_tctrain.json 2>$OUT/r2i_bench_tctrain.err; echo "bench tctrain: $?"
U2PL_ENTROPY_CHAIN=0 U2PL_TC_TRAIN=1 U2PL_CONV_STAGES=3 timeout 600 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2i_bench_tctrain_s3.json 2>$OUT/r2i_bench_tctrain_s3.err; echo "bench tctrain s3: $?"
tail -n 5 $OUT/r2i_pytest_entropy.log; tail -3 $OUT/r2i_chain_time.txt; cat $OUT/r2i_chain_time_clean.txt $OUT/r2i_chain_time_c19.txt | tail -2
python - <<'PY'
import json
for f in ['r2i_bench_tctrain','r2i_bench_tctrain_s3']:
    try:
        d=json.load(open(f'gpurun_out/{f}.json')); print(f, d['ms_per_step'], d['phases_ms'], d['losses'])
    except Exception as e: print(f, 'ERR', e)
PY
