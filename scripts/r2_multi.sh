#!/bin/bash
# multi-GPU check: C-ABI communicator tests + the bench under torchrun at N = $1
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_sharding_gloo.py -x -q > gpurun_out/m${N}_pytest.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
tail -4 gpurun_out/m${N}_pytest.log; tail -3 gpurun_out/m${N}_bench.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/m${N}_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d['e2e']); print(d['detail']['tree_broadcast']); c5=d['detail']['config5']; print({k:c5[k] for k in c5 if k!='checksum128_per_array'})
except Exception as e: print('ERR',e); print(open('gpurun_out/m${N}_bench.json').read()[-1500:])
PY
