run() { # name, env..., extra args
  name=$1; shift
  out=$(env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu --krylov-iters 31 $EXTRA 2>/dev/null | tail -1)
  echo "$name $(echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],2), round(d['ms_per_step'],4), round(d['krylov']['ms_per_iter'],4))" 2>/dev/null)"
  PORT=$((PORT+1))
}
PORT=29600
EXTRA="--overlap 0"; run default_serial A=1
EXTRA="--overlap 0"; run LL128_serial NCCL_PROTO=LL128
EXTRA="--overlap 0"; run LL_serial NCCL_PROTO=LL
EXTRA="--overlap 0"; run nch2_serial NCCL_MAX_NCHANNELS=2
EXTRA="--overlap 2"; run nch2_overlap NCCL_MAX_NCHANNELS=2
EXTRA="--overlap 2"; run nch1_LL_overlap NCCL_MAX_NCHANNELS=1 NCCL_PROTO=LL
