#!/bin/bash
# round 2, call AI: 2-GPU weak-scaling line at the final source state (torchrun, one rank per GPU, no collective on the data path)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_ai_bench_2gpu.json 2> gpurun_out/r02_ai_bench_2gpu.err
tail -1 gpurun_out/r02_ai_bench_2gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('2 GPUs: %.2f M/s e2e %.2f frac %.4f n_gpus %d scaling %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['n_gpus'], d['scaling']))"
tail -3 gpurun_out/r02_ai_bench_2gpu.err
