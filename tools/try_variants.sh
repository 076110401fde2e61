# on the GPU box: time the library variants of variants/ on the given workloads: tools/try_variants.sh "A B C" "config2 config3"
cp soap_b200/libsoap_b200.so /tmp/lib_orig.so
for v in $1; do
  cp variants/lib_$v.so soap_b200/libsoap_b200.so
  for w in $2; do
    timeout 200 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v','$w',round(d['ms_per_step'],2),'serial',d['serial_ms_per_step'],'ok',d['config']['halos_ok'],'rounds',d['stats']['rounds'])"
  done
done
cp /tmp/lib_orig.so soap_b200/libsoap_b200.so
