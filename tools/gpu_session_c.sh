#!/bin/bash
# Short one-GPU session: parity suite, bench config 1, and ncu captures of the two slow spots (legacy MiSeq replay, deflate).
set -u
TAG=${1:-r2f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -4 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
( timeout 600 python bench.py --steps 5 --warmup 3 ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
python - <<PY | tee -a $OUT/summary.txt
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n1.json").read().splitlines() if l.startswith("{")][-1])
    print("cfg1: value %.4g e2e %.4g gz %.4g (ratio %.2f) launch_ms %.3f" % (d["value"], d["e2e"]["value"], d["e2e_gzip"]["value"], d["e2e_gzip"]["compression_ratio"], d["roofline"]["avg_launch_ms"]))
except Exception as ex:
    print("parse failed", ex)
PY
( timeout 200 python tools/profile_model.py MiSeq 1000000 ) > $OUT/legacy_run.log 2>&1; tail -6 $OUT/legacy_run.log | tee -a $OUT/summary.txt
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:iss_generate_kernel -s 2 -c 1 -o $OUT/legacy_full -f python tools/profile_model.py MiSeq 1000000 ) > $OUT/ncu_legacy.log 2>&1
ncu -i $OUT/legacy_full.ncu-rep --page details > $OUT/legacy_ncu_details.txt 2>&1
ISS_KERNEL_THREADS=512 python tools/ncu_lines.py $OUT/legacy_full.ncu-rep 60 > $OUT/legacy_by_source_line.txt 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:iss_deflate_kernel -s 4 -c 1 -o $OUT/deflate_full -f python tools/gz_profile.py ) > $OUT/ncu_deflate.log 2>&1
ncu -i $OUT/deflate_full.ncu-rep --page details > $OUT/deflate_ncu_details.txt 2>&1
ISS_KERNEL_FILE=iss_deflate ISS_KERNEL_SYM=_Z18iss_deflate_kernel python tools/ncu_lines.py $OUT/deflate_full.ncu-rep 50 > $OUT/deflate_by_source_line.txt 2>&1
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"iss_gz|iss_deflate" -c 36 --csv --log-file $OUT/gzip_launches.csv python tools/gz_profile.py ) > $OUT/gz_profile.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:iss_generate_kernel -s 6 -c 1 -o $OUT/gen_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --device-only ) > $OUT/ncu_full.log 2>&1
ncu -i $OUT/gen_full.ncu-rep --page details > $OUT/ncu_details.txt 2>&1
ncu -i $OUT/gen_full.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>&1
python tools/ncu_lines.py $OUT/gen_full.ncu-rep 60 > $OUT/ncu_by_source_line.txt 2>&1
cat $OUT/summary.txt
