#!/bin/bash
# 2-GPU validation of bench.py (strong scaling, side-stream all-reduce per ELBO evaluation) + NCCL test of the sharded ELBO
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call20.log
: > $LOG
nvidia-smi -L | tee -a $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench_n2.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2.json'))
print('N=2 value %.3e ms/step %.4f kernel_us %.2f scaling %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_us'], d['scaling']))
print(' config', d['config'])
print(' e2e', d['e2e'])
print(' also', json.dumps(d['also'])[:700])
" 2>&1 | tee -a $LOG
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_bench_n1.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1.json'))
print('N=1 value %.3e ms/step %.4f kernel_us %.2f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_us']))
print(' also', json.dumps(d['also'])[:300])
" 2>&1 | tee -a $LOG
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-300 | tee -a $LOG
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3 | tee -a $LOG
exit 0
