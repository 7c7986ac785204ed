#!/bin/bash
# Round 2, 2-GPU call (gpurun --gpus 2): real GpuBank shards + the C-ABI NCCL group against the oracle; bench at N = 2.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/mg2_gpus.txt
timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/mg2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/mg2_pytest.log; tail -12 gpurun_out/mg2_pytest.log
tr() { local name=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus 2 --steps 5 --warmup 3 "$@" > gpurun_out/mg2_$name.json 2> gpurun_out/mg2_$name.err; tail -c 400 gpurun_out/mg2_$name.json; echo; }
tr saw_svf
tr subtractive_weak --workload subtractive
tr subtractive_strong --workload subtractive --scaling strong
tr net_strong --workload net --scaling strong
