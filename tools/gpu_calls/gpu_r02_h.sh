#!/bin/bash
# Round 2, GPU call H (8 GPUs): the bench as the driver launches it, and cfg5 with one neighbourhood shard per GPU.
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02h_bench8.json 2> gpurun_out/r02h_bench8.err; echo "bench8 rc=$?"; tail -1 gpurun_out/r02h_bench8.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e_dos_median']['value']); print('sharded_call', d.get('sharded_call')); print('gather', d.get('rank_sharded_gather')); print({k:(round(v['value'],1), round(v.get('e2e_dos_median',v.get('e2e'))['value'],1)) for k,v in d['workloads'].items()})"; tail -3 gpurun_out/r02h_bench8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload cfg5 --taxa 500 --bp 1500 --spr-rounds 3 --skip-cpu > gpurun_out/r02h_cfg5_8.json 2> gpurun_out/r02h_cfg5_8.err; echo "cfg5x8 rc=$?"; tail -1 gpurun_out/r02h_cfg5_8.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['tree'])"; tail -3 gpurun_out/r02h_cfg5_8.err
timeout 300 python tools/pcie_probe.py > gpurun_out/r02h_pcie1.json 2>&1; tail -2 gpurun_out/r02h_pcie1.json
