# compute-sanitizer (memcheck + racecheck) over small invocations that reach every kernel family of the halo path:
# staged tiers (DMO and hydro, kappa), general path incl. cluster scan, projected apertures, iterative tensors, mesh API
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import numpy as np, sys
from soap_b200 import synth
from soap_b200.halo_tasks import DeviceChunk, process_halos
from soap_b200.shared_mesh import SharedMesh
from tests import _compare as cmp
L = 30.0
cp = synth.coordinate_unit_params(L)
SO4 = [("crit", 200.0), ("mean", 200.0), ("crit", 500.0), ("BN98", float(synth.virBN98()))]
# DMO: tiers + general path (one halo above 32768 records -> cluster scan)
data, H = synth.to_numpy(*synth.nfw_chunk(160000, 120, L, seed=11, max_np=60000))
ch = DeviceChunk(data, L)
r = process_halos(ch, cmp.device_config(cp, so=SO4, flags=8, dmo=True), H)
print("dmo ok", int((r.status.cpu().numpy() == 0).sum()), ch.last_pairs())
m = SharedMesh(None, data[1]["Coordinates"], 12)
m.query_radius_periodic(H["cofp"][0], H["search_radius"][0], None, L)
ch.free()
# hydro: apertures, kinematics, kappa, tensors (+ iterative), half-mass radii, projected
data, H = synth.to_numpy(*synth.nfw_chunk(120000, 100, L, seed=12, max_np=20000, type_fractions={0: .45, 1: .5, 4: .049, 5: .001}))
aps = [(kpc * 1e-3 * cp["phys_mpc_to_coord"], kpc * 1e-3, incl) for kpc in (30.0, 100.0) for incl in (0, 1)]
ch = DeviceChunk(data, L)
r = process_halos(ch, cmp.device_config(cp, so=SO4[:2], apertures=aps, flags=1 | 2 | 4 | 8, dmo=False), H)
print("hydro tiers ok", int((r.status.cpu().numpy() == 0).sum()))
r = process_halos(ch, cmp.device_config(cp, so=SO4[:1], apertures=aps[:2], flags=1 | 4 | 8 | 16, dmo=False,
                                        projected=[(0.03 * cp["phys_mpc_to_coord"], 0.03), (0.1 * cp["phys_mpc_to_coord"], 0.1)]), H)
print("hydro general + projected + iterative ok", int((r.status.cpu().numpy() == 0).sum()))
ch.free()
PY
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/sanitizer_$tool.log
done
