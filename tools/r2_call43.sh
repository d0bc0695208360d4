#!/bin/bash
# round-2 call 43 (2 GPUs): the driver's N=2 launches of both arms, the NCCL test; M=64 accumulator layout probe
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call43.log
: > $LOG
timeout 60 ./tools/mma_mnmajor 2>&1 | tail -14 | tee -a $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench_n2_final.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_final.json'))
print('N=2 bench: value %.3e  ms/step %.4f  kernel_us %.2f  e2e %.3e  also %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_us'], d['e2e']['value'], json.dumps({k: {kk: vv for kk, vv in v.items() if kk in ('value','ms_per_step')} for k, v in (d.get('also') or {}).items()})))" 2>&1 | tee -a $LOG
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-250 | sed 's/^/[reference arm N=2] /' | tee -a $LOG
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_bench_n1_same_box_final.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_same_box_final.json'))
print('N=1 same box: value %.3e  kernel_us %.2f' % (d['value'], d['roofline']['kernel_us']))" 2>&1 | tee -a $LOG
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee -a $LOG
exit 0
