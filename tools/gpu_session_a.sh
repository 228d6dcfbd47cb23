#!/bin/bash
# One GPU session: parity suite, short bench (+ A/B switches of the quality-row ring), ncu launch list + full capture.
# Usage (from the repo root on the GPU box):  bash tools/gpu_session_a.sh <tag>
set -u
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
( timeout 120 python __graft_entry__.py smoke ) > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
( timeout 600 python bench.py --steps 5 --warmup 3 ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
python - <<PY | tee -a $OUT/summary.txt
import json
try:
    d=json.loads(open("$OUT/bench_n1.json").read().strip().splitlines()[-1])
    print("value %.4g e2e %.4g gz %.4g frac %.4f launch_ms %.3f" % (d["value"], d["e2e"]["value"], d["e2e_gzip"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"]))
except Exception as e:
    print("bench parse failed", e)
PY
( timeout 120 python tools/d2h_probe.py --seconds 0.7 --out $OUT/d2h_probe_n1.json ) > $OUT/d2h_probe_n1.log 2>&1; echo "probe rc=$?" | tee -a $OUT/summary.txt
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > $OUT/ncu_list.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:iss_generate_kernel -s 6 -c 1 -o $OUT/gen_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --device-only ) > $OUT/ncu_full.log 2>&1
ncu -i $OUT/gen_full.ncu-rep --page details > $OUT/ncu_details.txt 2>&1
ncu -i $OUT/gen_full.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>&1
python tools/ncu_lines.py $OUT/gen_full.ncu-rep 50 > $OUT/ncu_by_source_line.txt 2>&1
( timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "golden_fixture and (stress or miseq150 or hiseq_multi or read_length)" ) > $OUT/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/sanitizer_memcheck.txt | tee -a $OUT/summary.txt
cat $OUT/summary.txt
