#!/bin/bash
# round 2, call AL: 8-GPU weak-scaling line at the final source state (torchrun, one rank per GPU)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_al_bench_8gpu.json 2> gpurun_out/r02_al_bench_8gpu.err
tail -1 gpurun_out/r02_al_bench_8gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('8 GPUs: %.2f M/s e2e %.2f frac %.4f n_gpus %d scaling %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['n_gpus'], d['scaling']))"
