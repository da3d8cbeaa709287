#!/bin/bash
# 2-GPU run: NCCL sharding test + bench at N=1 and N=2 (weak scaling)
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu -rA > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee gpurun_out/summary.txt
tail -5 gpurun_out/pytest_multi.log
timeout 300 python bench.py --gpus 1 --steps 30 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale.err
cat gpurun_out/scale_n1.json | cut -c1-250
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/scale_n2.json 2>> gpurun_out/scale.err
echo "bench n2 rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/scale_n2.json | cut -c1-400; tail -5 gpurun_out/scale.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/scale_ref_n2.json 2>> gpurun_out/scale.err
cat gpurun_out/scale_ref_n2.json | cut -c1-300
