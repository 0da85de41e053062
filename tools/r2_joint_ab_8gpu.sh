set -u
best() { SWC_SHARD_JOINT=$1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2955$1 tools/api_bench.py --reps 4 --long-items 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['wall_s'], d['all_ms'], d['sharded_equals_single_gpu'], file=sys.stderr); print(d['wall_s'])"; }
J1=$(best 1); J0=$(best 0)
echo "joint=1 $J1  joint=0 $J0" | tee gpurun_out/r2_joint_ab_8gpu.txt
J=$(python -c "print(1 if $J1 <= $J0 else 0)")
echo "bench with SWC_SHARD_JOINT=$J" | tee -a gpurun_out/r2_joint_ab_8gpu.txt
SWC_SHARD_JOINT=$J python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_8gpu_j.json 2> gpurun_out/r2_bench_8gpu_j.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_8gpu_j.json'));print(d['value'],d['e2e']['value'],d['ms_per_step']);print({k:(v['value'],v['wall_s'],v['value_pageable_inputs'],v['equals_single_gpu']) for k,v in d['sharded_api'].items() if k.startswith('conf')})"
