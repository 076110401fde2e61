# round-1 capture of a launch window of one step (superseded by tools/ncu_kernels.sh, kept for the provenance of profiles/r01_ncu_full_final.txt)
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:^k_ --launch-skip 60 -c 70 -o /tmp/full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la /tmp/full.ncu-rep
python profiles/summarize_ncu.py report /tmp/full.ncu-rep > gpurun_out/ncu_full_final_summary.txt
tail -3 gpurun_out/ncu_full_final_summary.txt
grep -c "^\[" gpurun_out/ncu_full_final_summary.txt
