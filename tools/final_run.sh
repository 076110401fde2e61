# end-of-round measurement: default bench (config 2) + hydro variants; outputs under gpurun_out/
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 300 gpurun_out/bench_final.json
for w in config3 config3_kappa config3_iter; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/bench_final_$w.json
  python -c "import json; d=json.load(open('gpurun_out/bench_final_$w.json')); print('$w', d['ms_per_step'], d['value'])"
done
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_final_reference.json
cut -c1-400 gpurun_out/bench_final_reference.json
