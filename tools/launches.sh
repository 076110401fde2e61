# ncu launch list (device time per launch) of one bench step: tools/launches.sh <tag> [bench args]
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e "$@" > gpurun_out/launches_$tag.log 2>&1
python tools/sum_launches.py gpurun_out/launches_$tag.csv > gpurun_out/launches_$tag.txt
cat gpurun_out/launches_$tag.txt
