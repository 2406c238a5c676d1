set -u
ncu --set full --clock-control none --import-source on -k "regex:syncmers_fast" --launch-skip 3 -c 1 -f -o gpurun_out/prof_syn python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/prof_syn.ncu-rep
