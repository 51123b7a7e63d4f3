#!/bin/bash
# Round 2, final 8-GPU call: the bench as the driver launches it, final build.
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02final_bench8.json 2> gpurun_out/r02final_bench8.err; echo "bench8 rc=$?"; tail -1 gpurun_out/r02final_bench8.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e_dos_median']['value']); print('sharded_call', d.get('sharded_call',{}).get('value')); print({k:(round(v['value'],1), v.get('unit','GCUPS')) for k,v in d['workloads'].items()})"; tail -3 gpurun_out/r02final_bench8.err
