# (the .ncu-rep stays in /tmp on the box: with source pages it exceeds what gpurun copies back)
# ncu --set full of selected kernels of one bench step: tools/ncu_kernels.sh <tag> <regex> <count> [bench args]
tag=$1; rx=$2; cnt=$3; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$rx -c $cnt -o /tmp/ncu_$tag -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_$tag.log 2>&1
python profiles/summarize_ncu.py report /tmp/ncu_$tag.ncu-rep > gpurun_out/ncu_$tag.txt 2>&1
tail -5 gpurun_out/ncu_$tag.txt
