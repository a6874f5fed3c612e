set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
run 900 "pytest gpu" python -m pytest tests -m gpu -q > $OUT/r2v_pytest_gpu.log 2>&1
tail -4 $OUT/r2v_pytest_gpu.log
run 900 "bench full" python bench.py --steps 10 --warmup 3 --phases > $OUT/r2v_bench_full.json 2>$OUT/r2v_bench_full.err
run 400 "bench c2" python bench.py --workload c2 --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2v_bench_c2.json 2>$OUT/r2v_bench_c2.err
U2PL_CHAIN_TIMING=1 run 60 "chain dbg1" python tools/chain_time.py > $OUT/r2v_chain_time_dbg1.txt 2>&1
grep "entropy_chain. " $OUT/r2v_chain_time_dbg1.txt | tail -3
C=19 run 60 "chain c19" python tools/chain_time.py > $OUT/r2v_chain_time_c19.txt 2>&1
run 200 "ncu chain" ncu --set full --clock-control none --import-source on -k regex:entropy_chain -s 2 -c 1 -o $OUT/r02_entropy_chain_final python tools/chain_time.py > $OUT/r2v_ncu_chain.log 2>&1
run 100 "contra_bench" python tools/contra_bench.py > $OUT/r2v_contra_bench.txt 2>&1
run 120 "selftest perf" ./tools/cu/tc_selftest.bin perf > $OUT/r2v_perf_flat.txt 2>&1
U2PL_BENCH_FAST=1 run 900 "ncu launch list" ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r02_bench_step_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline > $OUT/r2v_bench_under_ncu.log 2>&1
python - <<'PY'
import json
for f in ['r2v_bench_full','r2v_bench_c2']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').readline()); print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['phases_ms'], d['losses'], d['roofline']['us_per_call'], d['roofline']['frac'], d['tensor_roofline']['frac'], d.get('eager_baseline'), d.get('cpu_baseline'), d['gpu_launches'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
