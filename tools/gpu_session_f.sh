#!/bin/bash
# Lean one-GPU session after a kernel change: smoke, parity suite, bench configs[1], launch list, ncu --set full capture of the
# generate kernel (details / raw / by source line).  Usage: bash tools/gpu_session_f.sh <tag>; then tools/collect_profiles.py.
set -u
TAG=${1:-r2s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 120 python __graft_entry__.py smoke ) > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -4 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
( timeout 600 python bench.py --steps 5 --warmup 3 ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
python - <<PY | tee -a $OUT/summary.txt
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n1.json").read().splitlines() if l.startswith("{")][-1])
    e=d.get("e2e",{}); z=d.get("e2e_gzip",{}); f=d.get("e2e_files",{})
    print("cfg1: value %.4g e2e %.4g gz %.4g (ratio %.2f) files %s launch_ms %.3f frac %.4f traffic %s" % (d["value"], e.get("value",0), z.get("value",0), z.get("compression_ratio",0), f.get("value"), d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"].get("traffic")))
except Exception as ex:
    print("parse failed", ex)
PY
for c in 2 3; do
( timeout 300 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --device-only ) > $OUT/bench_cfg${c}_dev.json 2> $OUT/bench_cfg${c}_dev.err
python - <<PY | tee -a $OUT/summary.txt
import json
try:
    d=json.loads([l for l in open("$OUT/bench_cfg${c}_dev.json").read().splitlines() if l.startswith("{")][-1])
    print("cfg$c device-only: value %.4g launch_ms %.3f" % (d["value"], d["roofline"]["avg_launch_ms"]))
except Exception as ex:
    print("parse failed", ex)
PY
done
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > $OUT/ncu_list.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:iss_generate_kernel -s 6 -c 1 -o $OUT/gen_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --device-only ) > $OUT/ncu_full.log 2>&1
ncu -i $OUT/gen_full.ncu-rep --page details > $OUT/ncu_details.txt 2>&1
ncu -i $OUT/gen_full.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>&1
ISS_KERNEL_SYM=${GEN_SYM:-_Z19iss_generate_kernelILi512ELi2ELi2ELi64E} python tools/ncu_lines.py $OUT/gen_full.ncu-rep 60 > $OUT/ncu_by_source_line.txt 2>&1
cat $OUT/summary.txt
