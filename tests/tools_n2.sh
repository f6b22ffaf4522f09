#!/bin/bash
# the driver's N=2 launch of both arms
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 3 2> gpurun_out/n2.err | tail -1 > gpurun_out/r02_bench_n2.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n2.json')); print({k: d.get(k) for k in ('value','n_gpus','ms_per_step','parity_vs_reference')}); print('encode',d['encode']['value'],'decode',d['decode']['value']); print('e2e',d['e2e']['value'],d['e2e']['encode']['value'],d['e2e']['decode']['value'])"
tail -3 gpurun_out/n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
