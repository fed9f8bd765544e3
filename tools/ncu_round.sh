# ncu evidence for bench.py's default workload (one GPU). Usage: bash tools/ncu_round.sh <tag>
# Each ncu pass follows a plain run of the same command that exited 0.
TAG=${1:-r02}
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_tb|k_bulk' -s 5 -c 3 -f -o gpurun_out/${TAG}_bulk $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture (A-B bulk) rc=$?"



ls -la gpurun_out/ | grep ${TAG}
