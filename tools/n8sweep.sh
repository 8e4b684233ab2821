mkdir -p gpurun_out
N=${1:-8}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r1_n${N}_$name.json 2> gpurun_out/r1_n${N}_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/r1_n${N}_$name.json')); print('$name', d['n_gpus'], round(d['value']), round(d['ms_per_step'],1), round(d['roofline']['achieved']), d['clocks']['sm_mhz'], d['final_loss'], round(d['e2e']['value']))" || tail -5 gpurun_out/r1_n${N}_$name.err
}
run peer B200_PEER_COMM=1
run nccl B200_PEER_COMM=0
