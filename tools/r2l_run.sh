set -u
OUT=gpurun_out
mkdir -p $OUT
U2PL_CHAIN_TIMING=1 timeout 120 python tools/chain_time.py > $OUT/r2l_chain_time_dbg1.txt 2>&1; echo "chain dbg1: $?"
timeout 120 python tools/chain_time.py > $OUT/r2l_chain_time_clean.txt 2>&1; echo "chain clean: $?"
timeout 300 python -m pytest tests/test_gpu_entropy.py -q > $OUT/r2l_pytest_entropy.log 2>&1; echo "pytest entropy: $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:entropy_chain -s 2 -c 1 -o $OUT/r02_entropy_chain_v4 python tools/chain_time.py > $OUT/r2l_ncu_chain.log 2>&1; echo "ncu chain: $?"
U2PL_ENTROPY_CHAIN=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:entropy_fast_hist -s 2 -c 1 -o $OUT/r02_entropy_fast_hist python tools/chain_time.py > $OUT/r2l_ncu_fasthist.log 2>&1; echo "ncu fast_hist: $?"
U2PL_TC_CONV=1 timeout 600 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2l_bench_tcconv1.json 2>$OUT/r2l_bench_tcconv1.err; echo "bench tcconv=1: $?"
timeout 900 python bench.py --steps 10 --warmup 3 --phases > $OUT/r2l_bench_full.json 2>$OUT/r2l_bench_full.err; echo "bench full: $?"
grep entropy_chain $OUT/r2l_chain_time_dbg1.txt | tail -2; cat $OUT/r2l_chain_time_clean.txt | tail -1; tail -2 $OUT/r2l_pytest_entropy.log
python - <<'PY'
import json
for f in ['r2l_bench_tcconv1','r2l_bench_full']:
    try:
        d=json.load(open(f'gpurun_out/{f}.json')); print(f, d['ms_per_step'], d['phases_ms'], d['losses'], d['roofline']['us_per_call'], d.get('eager_baseline'), d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
