set -x
timeout 300 python -m pytest tests/test_parallel_gpu.py -q -m gpu 2>&1 | tail -3
run() { PG_NCCL_MAX_CTAS=$1 PG_DP_RESERVE_SMS=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas=$1 reserve=$2', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['e2e']['value'])"; }
run 0 0
run 4 0
run 4 4
run 8 8
run 2 2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1 gpu', d['value'], d['ms_per_step'], d['roofline']['achieved'])"
