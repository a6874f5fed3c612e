set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 120 ./tools/cu/tc_selftest.bin all > $OUT/r2j_selftest_all.txt 2>&1; echo "selftest all: $?"
U2PL_CONV_STAGES=3 timeout 120 ./tools/cu/tc_selftest.bin conv > $OUT/r2j_selftest_conv_s3.txt 2>&1; echo "selftest conv s3: $?"
U2PL_CONV_STAGES=4 timeout 120 ./tools/cu/tc_selftest.bin conv > $OUT/r2j_selftest_conv_s4.txt 2>&1; echo "selftest conv s4: $?"
U2PL_CONV_STAGES=3 timeout 120 ./tools/cu/tc_selftest.bin stats > $OUT/r2j_selftest_stats_s3.txt 2>&1; echo "selftest stats s3: $?"
U2PL_CONV_STAGES=4 timeout 120 ./tools/cu/tc_selftest.bin stats > $OUT/r2j_selftest_stats_s4.txt 2>&1; echo "selftest stats s4: $?"
timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2j_perf_flat.txt 2>&1; echo "perf flat: $?"
U2PL_CONV_STAGES=3 timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2j_perf_flat_s3.txt 2>&1; echo "perf flat s3: $?"
U2PL_CONV_STAGES=4 timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2j_perf_flat_s4.txt 2>&1; echo "perf flat s4: $?"
U2PL_CHAIN_TIMING=1 timeout 120 python tools/chain_time.py > $OUT/r2j_chain_time.txt 2>&1; echo "chain timing: $?"
timeout 120 python tools/chain_time.py > $OUT/r2j_chain_time_clean.txt 2>&1; echo "chain clean: $?"
C=19 timeout 120 python tools/chain_time.py > $OUT/r2j_chain_time_c19.txt 2>&1; echo "chain c19: $?"
timeout 900 python -m pytest tests -m gpu -q > $OUT/r2j_pytest_gpu.log 2>&1; echo "pytest all: $?"
U2PL_TC_TRAIN=1 timeout 600 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2j_bench_tctrain.json 2>$OUT/r2j_bench_tctrain.err; echo "bench tctrain: $?"
timeout 600 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2j_bench.json 2>$OUT/r2j_bench.err; echo "bench: $?"
grep -v OK $OUT/r2j_selftest_*.txt; cat $OUT/r2j_perf_flat.txt; grep "1x1" $OUT/r2j_perf_flat_s3.txt $OUT/r2j_perf_flat_s4.txt; tail -n 8 $OUT/r2j_pytest_gpu.log; tail -3 $OUT/r2j_chain_time.txt; cat $OUT/r2j_chain_time_clean.txt $OUT/r2j_chain_time_c19.txt | tail -2
python - <<'PY'
import json
for f in ['r2j_bench_tctrain','r2j_bench']:
    try:
        d=json.load(open(f'gpurun_out/{f}.json')); print(f, d['ms_per_step'], d['phases_ms'], d['losses'], d['roofline']['us_per_call'])
    except Exception as e: print(f, 'ERR', e)
PY
