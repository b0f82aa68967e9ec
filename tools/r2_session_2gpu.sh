#!/bin/bash
# Round-2 multi-GPU session (N = $1): NCCL / peer-memory tests, the bench line at N GPUs.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 2>&1 | tail -4
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 1000 --warmup 10 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "bench rc=$?"
grep -E "NCCL INFO (Connected|comm 0x|ncclCommInitRank|NVLS|Channel 00/)" gpurun_out/r02_bench_${N}gpu.err | head -6
grep -v "NCCL INFO" gpurun_out/r02_bench_${N}gpu.err | tail -5
python - <<PY
import json
for line in open("gpurun_out/r02_bench_${N}gpu.json"):
  line = line.strip()
  if line.startswith("{"):
    d = json.loads(line)
    print("value", d["value"], "us/step", d["ms_per_step"] * 1e3, "e2e", d["e2e"]["value"], d["e2e"].get("host_link"))
    print("resgld", d.get("resgld"))
    print("row_sharded", d.get("row_sharded_gradient"))
    print(d["clocks"])
PY
