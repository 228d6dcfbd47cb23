#!/bin/bash
# One-GPU A/B session: parity suite, bench config 1 (+ the 128-pair-tile variant, device only), legacy MiSeq, deflate
# launch list and ncu capture, ncu capture of the generate kernel.
set -u
TAG=${1:-r2i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -4 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
( timeout 600 python bench.py --steps 5 --warmup 3 ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
summ() { python - "$1" "$2" <<'PY' | tee -a $OUT/summary.txt
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    gz = d.get("e2e_gzip") or {}
    print("%s: value %.4g e2e %.4g gz %.4g (ratio %.2f) launch_ms %.3f" % (sys.argv[2], d["value"], (d.get("e2e") or {}).get("value", 0), gz.get("value", 0), gz.get("compression_ratio", 0), d["roofline"]["avg_launch_ms"]))
except Exception as ex:
    print(sys.argv[2], "parse failed", ex)
PY
}
export OUT
summ $OUT/bench_n1.json cfg1
for t in 64 128; do
( ISS_TILE=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --device-only ) > $OUT/bench_tile$t.json 2> $OUT/bench_tile$t.err
summ $OUT/bench_tile$t.json tile$t
done
( timeout 300 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --device-only ) > $OUT/bench_cfg3_dev.json 2> $OUT/bench_cfg3_dev.err
summ $OUT/bench_cfg3_dev.json cfg3dev
( timeout 400 python bench.py --config 2 --pairs 20000000 --steps 3 --warmup 3 --no-cpu-baseline ) > $OUT/bench_cfg2_20M.json 2> $OUT/bench_cfg2_20M.err
summ $OUT/bench_cfg2_20M.json cfg2_20M
( timeout 300 python tools/fasta_load_bench.py 330 3000000 ) > $OUT/fasta_load.json 2> $OUT/fasta_load.err; tail -1 $OUT/fasta_load.json | tee -a $OUT/summary.txt
( timeout 200 python tools/profile_model.py MiSeq 1000000 ) > $OUT/legacy_run.log 2>&1; tail -3 $OUT/legacy_run.log | head -2 | tee -a $OUT/summary.txt
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"iss_gz|iss_deflate" -c 36 --csv --log-file $OUT/gzip_launches.csv python tools/gz_profile.py ) > $OUT/gz_profile.log 2>&1
python - <<'PY' | tee -a $OUT/summary.txt
import csv, collections, os
d = collections.defaultdict(list)
for r in csv.reader(open(os.environ["OUT"] + "/gzip_launches.csv")):
    if len(r) > 5 and r[0].isdigit():
        d[r[4].split("(")[0]].append(float(r[-1].replace(",", "")))
for k, v in d.items():
    print("%s: %d launches, mean %.1f us" % (k, len(v), sum(v) / len(v) / 1000))
PY
( timeout 120 python tools/gz_trace.py ) > $OUT/gz_trace.out 2> $OUT/gz_trace.log
grep device_only $OUT/gz_trace.log | tee -a $OUT/summary.txt
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:iss_deflate_kernel -s 4 -c 1 -o $OUT/deflate_full -f python tools/gz_profile.py ) > $OUT/ncu_deflate.log 2>&1
ncu -i $OUT/deflate_full.ncu-rep --page details > $OUT/deflate_ncu_details.txt 2>&1
ISS_KERNEL_FILE=iss_deflate ISS_KERNEL_SYM=_Z18iss_deflate_kernel python tools/ncu_lines.py $OUT/deflate_full.ncu-rep 50 > $OUT/deflate_by_source_line.txt 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:iss_generate_kernel -s 6 -c 1 -o $OUT/gen_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --device-only ) > $OUT/ncu_full.log 2>&1
ncu -i $OUT/gen_full.ncu-rep --page details > $OUT/ncu_details.txt 2>&1
ncu -i $OUT/gen_full.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>&1
ISS_KERNEL_SYM=${GEN_SYM:-_Z19iss_generate_kernelILi512ELi2ELi2ELi64E} python tools/ncu_lines.py $OUT/gen_full.ncu-rep 60 > $OUT/ncu_by_source_line.txt 2>&1
cat $OUT/summary.txt
