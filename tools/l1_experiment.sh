set -e
ISS_NVCC_EXTRA="-DISS_MIN_BLOCKS=2" python -c "from insilicoseq_b200._build import build_native; build_native(force=True)" >/dev/null 2>&1
ncu --set full --clock-control none -k regex:iss_generate -s 3 -c 1 -o gpurun_out/prof_mb2 python bench.py --steps 1 --warmup 3 --pairs 2500000 --no-cpu-baseline > /dev/null 2>&1
python -c "from insilicoseq_b200._build import build_native; build_native(force=True)" >/dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:iss_generate -s 3 -c 1 -o gpurun_out/prof_r1e python bench.py --steps 1 --warmup 3 --pairs 2500000 --no-cpu-baseline > /dev/null 2>&1
