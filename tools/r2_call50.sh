#!/bin/bash
# round-2 call 50 (2 GPUs): the driver's N=2 launch of bench.py on the final tree (after the one-graph refactor of the N=1 path)
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call50.log
: > $LOG
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench_n2_final2.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_final2.json'))
print('N=2 bench: value %.3e  ms/step %.4f  kernel_us %.2f  e2e %.3e  also %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_us'], d['e2e']['value'], json.dumps({k: {kk: vv for kk, vv in v.items() if kk in ('value','ms_per_step')} for k, v in (d.get('also') or {}).items()})))" 2>&1 | tee -a $LOG
exit 0
