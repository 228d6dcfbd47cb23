set -e
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"$1\", round(d[\"value\"]/1e6,1),\"Mpairs/s\", round(d[\"roofline\"][\"avg_launch_ms\"],3),\"ms/launch\")" | tee -a gpurun_out/variants.txt; }
python -m pytest tests -q -m gpu -x 2>&1 | tail -2
run "minblocks4"
ISS_NVCC_EXTRA="-DISS_MIN_BLOCKS=3" python -c "from insilicoseq_b200._build import build_native; build_native(force=True)" >/dev/null 2>&1
run "minblocks3"
ISS_NVCC_EXTRA="-DISS_MIN_BLOCKS=2" python -c "from insilicoseq_b200._build import build_native; build_native(force=True)" >/dev/null 2>&1
run "minblocks2"
