for n in 125000 250000 500000 1000000; do
  python tools/tune_probe.py $n resident
  PM_COUNT_WARP_BELOW=100000000 python tools/tune_probe.py $n resident
done
python tools/tune_probe.py 1000000 e2e
PM_COUNT_WARP_BELOW=300000 python tools/tune_probe.py 1000000 e2e
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_e2e_packed.csv python tools/tune_probe.py 1000000 e2e > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_e2e_packed.csv | head -40
