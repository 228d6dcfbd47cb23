#!/bin/bash
# Two-GPU sanity session for the closing build: the multi-GPU CLI tests (NCCL broadcast of the genome loaded from text,
# GPU-count invariance up to the GPUs present) and one torchrun bench line at N = 2.  Usage: bash tools/gpu_session_2.sh <tag>
set -u
TAG=${1:-r2n2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
N=$(nvidia-smi -L | wc -l)
echo "gpus=$N" | tee $OUT/summary.txt
( timeout 600 python -m pytest tests/test_gpu_cli.py -q -x ) > $OUT/pytest_cli.log 2>&1; echo "pytest cli rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/pytest_cli.log | tee -a $OUT/summary.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline ) > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
python - <<PY | tee -a $OUT/summary.txt
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n$N.json").read().splitlines() if l.startswith("{")][-1])
    e=d.get("e2e",{}); z=d.get("e2e_gzip",{}); f=d.get("e2e_files",{})
    print("N=%d: value %.4g e2e %.4g gz %.4g files %s" % (d["n_gpus"], d["value"], e.get("value",0), z.get("value",0), f.get("value")))
except Exception as ex:
    print("parse failed", ex)
PY
cat $OUT/summary.txt
