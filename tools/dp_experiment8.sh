# 8-GPU bucket-granularity check (gpurun --gpus 8): transformer blocks per all-reduce
export NG=8
run() { PG_NCCL_MAX_CTAS=$1 PG_DP_RESERVE_SMS=$2 PG_DP_BUCKET_BLOCKS=$3 PG_DP_OVERLAP=$4 timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('gpus=$NG ctas=$1 reserve=$2 bucket_blocks=$3 overlap=$4:', d['value'], 'img/s', d['ms_per_step'], 'ms  GEMM', d['roofline']['achieved'], 'TF  e2e', d['e2e']['value'])"; }
run 0 0 24 1
run 0 0 6 1
run 0 0 1 1
