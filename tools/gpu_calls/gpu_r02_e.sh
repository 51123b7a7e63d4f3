#!/bin/bash
# Round 2, GPU call E (2 GPUs): the bench under torchrun, both multi-GPU paths, cfg5 sharded.
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02e_bench2.json 2> gpurun_out/r02e_bench2.err; echo "bench2 rc=$?"; tail -1 gpurun_out/r02e_bench2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e_dos_median']['value']); print('sharded_call', d.get('sharded_call')); print('gather', d.get('rank_sharded_gather')); print({k:(round(v['value'],1), round(v.get('e2e_dos_median',v.get('e2e'))['value'],1)) for k,v in d['workloads'].items()})"; tail -5 gpurun_out/r02e_bench2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg5 --taxa 500 --bp 1500 --spr-rounds 3 --skip-cpu > gpurun_out/r02e_cfg5_2.json 2> gpurun_out/r02e_cfg5_2.err; echo "cfg5x2 rc=$?"; tail -1 gpurun_out/r02e_cfg5_2.json | cut -c1-1600; tail -3 gpurun_out/r02e_cfg5_2.err
timeout 600 python bench.py --workload cfg5 --taxa 500 --bp 1500 --spr-rounds 3 --skip-cpu > gpurun_out/r02e_cfg5_1.json 2> gpurun_out/r02e_cfg5_1.err; echo "cfg5x1 rc=$?"; tail -1 gpurun_out/r02e_cfg5_1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['tree'])"
