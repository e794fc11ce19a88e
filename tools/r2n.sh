N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
e2e() { # tag env
  env $2 $TR bench.py --gpus $N --steps 6 --no-cpu --no-secondary 2>/dev/null | grep '^{' > gpurun_out/r2n_$1.json
  python -c "
import json; d=json.load(open('gpurun_out/r2n_$1.json')); print('$1', 'N', d['n_gpus'], 'e2e', round(d['e2e']['value'],2), 'per gpu', round(d['e2e']['value']/d['n_gpus'],2))"
}
e2e default A=1
e2e poll FLIC_WAIT=poll
e2e block FLIC_WAIT=block
e2e both4 FLIC_PIPE_BOTH_DEPTH=4
e2e both4poll "FLIC_PIPE_BOTH_DEPTH=4 FLIC_WAIT=poll"
