set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
run 300 "pytest 2gpu" python -m pytest tests/test_gpu_sharded_bank.py tests/test_gpu_syncbn.py -q -x > $OUT/r2p_pytest_2gpu.log 2>&1
tail -5 $OUT/r2p_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run 400 "bench n2 replicated+peer" $TR --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --phases > $OUT/r2p_bench_n2.json 2>$OUT/r2p_bench_n2.err
U2PL_BANK_SHARDED=1 run 400 "bench n2 sharded" $TR --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --phases > $OUT/r2p_bench_n2_sharded.json 2>$OUT/r2p_bench_n2_sharded.err
U2PL_PEER_SYNCBN=0 run 400 "bench n2 nccl syncbn" $TR --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 --phases > $OUT/r2p_bench_n2_ncclbn.json 2>$OUT/r2p_bench_n2_ncclbn.err
python - <<'PY'
import json
for f in ['r2p_bench_n2','r2p_bench_n2_sharded','r2p_bench_n2_ncclbn']:
    try:
        for line in open(f'gpurun_out/{f}.json'):
            if line.startswith('{'):
                d=json.loads(line); print(f, d['ms_per_step'], d['value'], d['phases_ms'], d['losses'], d['config']['bank']); break
        else: print(f, 'no json'); print(open(f'gpurun_out/{f}.err').read()[-1500:])
    except Exception as e: print(f, 'ERR', e)
PY
