# ncu --set full with source pages of one kernel (default: the tier-0 front; round 1 used k_small_warps): tools/ncu_tier0.sh [kernel regex]
rx=${1:-k_tier_front}
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -o /tmp/tier0 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tier0.log 2>&1
ls -la /tmp/tier0.ncu-rep
ncu -i /tmp/tier0.ncu-rep --page source --csv > gpurun_out/tier0_source.csv 2>/dev/null
ncu -i /tmp/tier0.ncu-rep --page details --csv > gpurun_out/tier0_details.csv 2>/dev/null
python profiles/summarize_ncu.py report /tmp/tier0.ncu-rep > gpurun_out/tier0_summary.txt
cat gpurun_out/tier0_summary.txt
ls -la gpurun_out
