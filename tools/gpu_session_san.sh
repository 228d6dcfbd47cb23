#!/bin/bash
# Sanitizer pass over the built library: memcheck on the generate kernel (all builds: dense indels, ties, list overflow, short
# rows, long reads), on the gzip kernels and on the genome-file kernels; racecheck on three generate cases.
set -u
TAG=${1:-r2x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "(golden_fixture and (stress or miseq150 or hiseq_multi or read_length or legacy or full_indel or ties or low_quality or long_reads)) or short_quality_rows" ) > $OUT/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/sanitizer_memcheck.txt | tee -a $OUT/summary.txt
( timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "gzip_name_line or compressed_batches or distance_one_runs or genome_file_as_text" ) > $OUT/sanitizer_gzip.txt 2>&1; echo "memcheck gzip rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/sanitizer_gzip.txt | tee -a $OUT/summary.txt
( timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "golden_fixture and (stress_indels_dense or miseq150 or novaseq_single)" ) > $OUT/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/sanitizer_racecheck.txt | tee -a $OUT/summary.txt
cat $OUT/summary.txt
