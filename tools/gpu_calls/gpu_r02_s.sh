#!/bin/bash
# Round 2, GPU call S (2 GPUs): the bench as the driver launches it on the final build (all workload blocks under torchrun), both arms.
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02s_bench2.json 2> gpurun_out/r02s_bench2.err; echo "bench2 rc=$?"; tail -1 gpurun_out/r02s_bench2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e_dos_median']['value']); print('sharded_call', d.get('sharded_call',{}).get('value')); print({k:(round(v['value'],1), v.get('unit','GCUPS')) for k,v in d['workloads'].items()})"; tail -3 gpurun_out/r02s_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02s_ref2.json 2> gpurun_out/r02s_ref2.err; echo "ref2 rc=$?"; cut -c1-300 gpurun_out/r02s_ref2.json
