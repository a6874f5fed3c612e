set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2b_pytest_gpu.log 2>&1; echo "pytest: $?"
timeout 400 python bench.py --steps 10 --warmup 3 --phases > $OUT/r2b_bench_n1.json 2>$OUT/r2b_bench_n1.err; echo "bench: $?"
U2PL_TC_WGRAD=1 timeout 300 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2b_bench_n1_tcwgrad.json 2>$OUT/r2b_bench_n1_tcwgrad.err; echo "bench tcwgrad: $?"
U2PL_TC_CONV=0 timeout 300 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2b_bench_n1_notcconv.json 2>$OUT/r2b_bench_n1_notcconv.err; echo "bench no tcconv: $?"
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --phases --no-cpu-baseline > $OUT/r2b_bench_c2_n1.json 2>$OUT/r2b_bench_c2_n1.err; echo "bench c2: $?"
timeout 300 python tools/step_profile.py > $OUT/r2b_step_profile.txt 2>$OUT/r2b_step_profile.err; echo "profile: $?"
tail -n 3 $OUT/r2b_pytest_gpu.log
