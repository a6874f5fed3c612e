set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
U2PL_CHAIN_TIMING=1 run 60 "chain dbg1" python tools/chain_time.py > $OUT/r2m_chain_time_dbg1.txt 2>&1
run 60 "chain clean" python tools/chain_time.py > $OUT/r2m_chain_time_clean.txt 2>&1
C=19 run 60 "chain c19" python tools/chain_time.py > $OUT/r2m_chain_time_c19.txt 2>&1
run 120 "pytest entropy" python -m pytest tests/test_gpu_entropy.py -q -x > $OUT/r2m_pytest_entropy.log 2>&1
grep entropy_chain $OUT/r2m_chain_time_dbg1.txt | tail -2; cat $OUT/r2m_chain_time_clean.txt | tail -1; tail -2 $OUT/r2m_pytest_entropy.log
run 200 "ncu chain" ncu --set full --clock-control none --import-source on -k regex:entropy_chain -s 2 -c 1 -o $OUT/r02_entropy_chain_v4 python tools/chain_time.py > $OUT/r2m_ncu_chain.log 2>&1
U2PL_TC_CONV=1 run 300 "bench tcconv=1" python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2m_bench_tcconv1.json 2>$OUT/r2m_bench_tcconv1.err
U2PL_TC_CONV=1 U2PL_TC_T2=1 run 300 "bench tcconv=1 t2" python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2m_bench_tcconv1_t2.json 2>$OUT/r2m_bench_tcconv1_t2.err
run 600 "bench full" python bench.py --steps 10 --warmup 3 --phases > $OUT/r2m_bench_full.json 2>$OUT/r2m_bench_full.err
python - <<'PY'
import json
for f in ['r2m_bench_tcconv1','r2m_bench_tcconv1_t2','r2m_bench_full']:
    try:
        d=json.load(open(f'gpurun_out/{f}.json')); print(f, d['ms_per_step'], d['phases_ms'], d['losses'], d['roofline']['us_per_call'], d.get('eager_baseline'), d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
