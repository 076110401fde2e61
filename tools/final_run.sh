# round-2 measurement set on one B200: outputs under gpurun_out/ (copied to profiles/ by hand)
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --checked 2>&1 | tail -4) > gpurun_out/r02_checked_build_tests.txt
grep -c SOAP_ASSERT gpurun_out/r02_checked_build_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_config2.json 2> gpurun_out/r02_bench_config2.err
tail -c 400 gpurun_out/r02_bench_config2.json
for w in config3 config3_kappa config3_iter config4; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/r02_bench_$w.err | tail -1 > gpurun_out/r02_bench_$w.json
  python -c "import json; d=json.load(open('gpurun_out/r02_bench_$w.json')); print('$w', d['ms_per_step'], d['value'])"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference.json
cut -c1-300 gpurun_out/r02_bench_reference.json
bash tools/launches.sh r02_config2 > /dev/null 2>&1
bash tools/ncu_kernels.sh r02_config2 "k_" 170 > /dev/null 2>&1
tail -3 gpurun_out/ncu_r02_config2.txt
