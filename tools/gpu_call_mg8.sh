#!/bin/bash
# Round 2, 8-GPU call (gpurun --gpus 8): the configurations the north-star names for multi-GPU — config 4 "1 -> 8 GPU" (weak: 1024 voices per
# GPU; strong: 1024 voices in all) and config 5 "65536 voices sharded 8 x B200, NCCL mix-down" (strong: 65536 in all) — at N = 1, 2, 4, 8.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/mg8_gpus.txt
one() { local name=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then timeout 150 python bench.py --gpus 1 --steps 5 --warmup 3 "$@" > gpurun_out/mg8_${name}_n$n.json 2> gpurun_out/mg8_${name}_n$n.err
  else timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $n --steps 5 --warmup 3 "$@" > gpurun_out/mg8_${name}_n$n.json 2> gpurun_out/mg8_${name}_n$n.err; fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/mg8_${name}_n$n.json").read().strip().splitlines()[-1])
    print("${name} N=$n value %.0f e2e %.0f ms %.3f voices/gpu %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["voices_per_gpu"]))
except Exception as e:
    print("${name} N=$n FAILED", e)
PY
}
# (N = 1 and N = 2 of the same command lines come from the 1- and 2-GPU calls: an 8-GPU box is charged 8 x its time)
for n in 4 8; do one subtractive_weak $n --workload subtractive; done
for n in 4 8; do one subtractive_strong $n --workload subtractive --scaling strong; done
for n in 4 8; do one net_strong $n --workload net --scaling strong; done
for n in 8; do one saw_svf $n; one net_weak $n --workload net; done
