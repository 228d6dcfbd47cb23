#!/bin/bash
# The round's closing one-GPU session: parity suite, bench configs 1-4, model sweep, genome-file loading, gzip timeline, ncu launch
# lists + full captures (generate and deflate kernels), sanitizers.
# Usage (repo root on the GPU box):  bash tools/gpu_session_e.sh <tag>
set -u
TAG=${1:-r2z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
( timeout 120 python __graft_entry__.py smoke ) > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -4 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
show () {
python - <<PY | tee -a $OUT/summary.txt
import json
try:
    d=json.loads([l for l in open("$1").read().splitlines() if l.startswith("{")][-1])
    e=d.get("e2e",{}); z=d.get("e2e_gzip",{}); f=d.get("e2e_files",{}); h=d.get("host",{})
    print("$2: value %.4g e2e %.4g gz %.4g (ratio %.2f) files %s gzfiles %s set_genome %s s launch_ms %.3f frac %.4f" % (d["value"], e.get("value",0), z.get("value",0), z.get("compression_ratio",0), f.get("value"), f.get("gzip_value"), h.get("set_genome_s"), d["roofline"]["avg_launch_ms"], d["roofline"]["frac"]))
except Exception as ex:
    print("$2 parse failed", ex)
PY
}
( timeout 600 python bench.py --steps 5 --warmup 3 ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
show $OUT/bench_n1.json cfg1
( timeout 300 python tools/model_sweep.py 4000000 ) > $OUT/model_sweep.log 2>&1; cp gpurun_out/model_sweep.json $OUT/model_sweep.json 2>/dev/null; cat $OUT/model_sweep.log | tee -a $OUT/summary.txt
for c in 2 3 4; do
  ( timeout 900 python bench.py --config $c --steps 3 --warmup 3 ) > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err; echo "bench cfg$c rc=$?" | tee -a $OUT/summary.txt
  show $OUT/bench_cfg$c.json cfg$c
done
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"iss_gz|iss_deflate" -c 36 --csv --log-file $OUT/gzip_launches.csv python tools/gz_profile.py ) > $OUT/gz_profile.log 2>&1
( timeout 120 python tools/gz_trace.py ) > $OUT/gz_trace.out 2> $OUT/gz_trace.log
grep device_only $OUT/gz_trace.log | tee -a $OUT/summary.txt
( timeout 300 python tools/fasta_load_bench.py 660 3000000 ) > $OUT/fasta_load.json 2> $OUT/fasta_load.err; tail -1 $OUT/fasta_load.json | tee -a $OUT/summary.txt
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:iss_deflate_kernel -s 4 -c 1 -o $OUT/deflate_full -f python tools/gz_profile.py ) > $OUT/ncu_deflate.log 2>&1
ncu -i $OUT/deflate_full.ncu-rep --page details > $OUT/deflate_ncu_details.txt 2>&1
ISS_KERNEL_FILE=iss_deflate ISS_KERNEL_SYM=_Z18iss_deflate_kernel python tools/ncu_lines.py $OUT/deflate_full.ncu-rep 50 > $OUT/deflate_by_source_line.txt 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > $OUT/ncu_list.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:iss_generate_kernel -s 6 -c 1 -o $OUT/gen_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --device-only ) > $OUT/ncu_full.log 2>&1
ncu -i $OUT/gen_full.ncu-rep --page details > $OUT/ncu_details.txt 2>&1
ncu -i $OUT/gen_full.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>&1
ISS_KERNEL_SYM=${GEN_SYM:-_Z19iss_generate_kernelILi512ELi2ELi2ELi64E} python tools/ncu_lines.py $OUT/gen_full.ncu-rep 60 > $OUT/ncu_by_source_line.txt 2>&1
( timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "golden_fixture and (stress or miseq150 or hiseq_multi or read_length or legacy or full_indel or ties or low_quality)" ) > $OUT/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/sanitizer_memcheck.txt | tee -a $OUT/summary.txt
( timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "golden_fixture and (stress_indels_dense or miseq150 or legacy)" ) > $OUT/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/sanitizer_racecheck.txt | tee -a $OUT/summary.txt
( timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "gzip_name_line or compressed_batches or distance_one_runs or genome_file_as_text" ) > $OUT/sanitizer_gzip.txt 2>&1; echo "memcheck gzip rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/sanitizer_gzip.txt | tee -a $OUT/summary.txt
cat $OUT/summary.txt
