set -x
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 400 gpurun_out/bench_final.json
python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/bench_final_config3.json
python bench.py --workload config3_kappa --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/bench_final_config3_kappa.json
python bench.py --workload config3_iter --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/iter.err | tail -1 > gpurun_out/bench_final_config3_iter.json
for f in gpurun_out/bench_final_config3*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', d['ms_per_step'], d['value'], d.get('phases_ms'))"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01_final2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final2.log 2>&1
python profiles/summarize_ncu.py launches gpurun_out/launches_r01_final2.csv > gpurun_out/launches_r01_final2.txt
head -20 gpurun_out/launches_r01_final2.txt
