#!/bin/bash
# The 8-GPU session (one box, 8 x B200): host-side D2H probe, GPU-count invariance, bench configs 1 / 3 / 4 at N = 8,
# and the e2e leg under the three ring allocations.  Usage: bash tools/gpu_session_8.sh <tag>
set -u
TAG=${1:-r2n8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
N=$(nvidia-smi -L | wc -l)
echo "gpus=$N" | tee $OUT/summary.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
( lscpu; numactl -H; cat /sys/fs/cgroup/cpu.max; free -g ) > $OUT/host.txt 2>&1
( timeout 300 python tools/d2h_probe.py --seconds 1.0 --out $OUT/d2h_probe.json ) > $OUT/d2h_probe.log 2>&1; echo "probe rc=$?" | tee -a $OUT/summary.txt
( timeout 600 python -m pytest tests/test_gpu_cli.py -q -x -k "invariance or multi" ) > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/pytest_multi.log | tee -a $OUT/summary.txt
run_bench () {  # name, extra env, args
  local name=$1; shift
  local envs=$1; shift
  ( env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" ) > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "bench $name rc=$?" | tee -a $OUT/summary.txt
  python - <<PY | tee -a $OUT/summary.txt
import json
try:
    d=json.loads([l for l in open("$OUT/bench_$name.json").read().splitlines() if l.startswith("{")][-1])
    e=d.get("e2e",{}); z=d.get("e2e_gzip",{}); f=d.get("e2e_files",{})
    print("$name: value %.4g e2e %.4g (%.1f GB/s/gpu) gz %.4g files %s" % (d["value"], e.get("value",0), e.get("d2h_GBps_per_gpu",0), z.get("value",0), f.get("value")))
except Exception as ex:
    print("$name parse failed", ex)
PY
}
run_bench cfg1_numa "ISS_RING_ALLOC=numa" --steps 5 --warmup 3
run_bench cfg1_pinned "ISS_RING_ALLOC=pinned" --steps 5 --warmup 3
run_bench cfg1_wc "ISS_RING_ALLOC=wc" --steps 5 --warmup 3
run_bench cfg3 "ISS_RING_ALLOC=numa" --config 3 --steps 3 --warmup 3
run_bench cfg4 "ISS_RING_ALLOC=numa" --config 4 --steps 3 --warmup 3
( ISS_RING_ALLOC=numa timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 5 --warmup 3 ) > $OUT/bench_cfg1_n4.json 2> $OUT/bench_cfg1_n4.err
cat $OUT/summary.txt
