set -e
python -m pytest tests -q -m gpu -x 2>&1 | tail -2
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"$1\", round(d[\"value\"]/1e6,1),\"Mpairs/s\", round(d[\"roofline\"][\"avg_launch_ms\"],3),\"ms/launch\", \"e2e\", round(d[\"e2e\"][\"value\"]/1e6,1))" | tee -a gpurun_out/ab_tma.txt; }
run "tma_gather"
ISS_NO_TMA_GATHER=1 run "ldg_gather"
run "tma_gather"
ISS_NO_TMA_GATHER=1 run "ldg_gather"
