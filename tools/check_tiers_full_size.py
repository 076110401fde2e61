import numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
from soap_b200 import synth
from tests import test_gpu_full_size as T
L=284.4; cp=synth.coordinate_unit_params(L)
data, halos = synth.nfw_chunk(512**3, 200000, L, seed=20261018, device="cuda", max_np=2.0e6)
a, sa, pa = T._process(data, halos, cp, L)
b, sb, pb = T._process(data, halos, cp, L, no_tiers=True)
print("status equal", np.array_equal(sa, sb), "pairs", pa, pb)
T._assert_same(a, b, T.INT_KEYS)
worst = 0.0
for k in a:
    x, y = a[k], b[k]
    sc = np.maximum(np.abs(y), 1e-30)
    d = np.nanmax(np.abs(x - y) / (sc + 1e-9*np.nanmax(np.abs(y))+1e-300))
    worst = max(worst, d)
print("tiers == general path on all 200000 halos; worst relative difference", worst)
