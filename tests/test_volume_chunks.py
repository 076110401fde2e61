"""One volume, many chunks (soap_b200/synth.py: volume_catalogue / volume_chunk): a chunk generated on its own
carries exactly the particles of the whole volume that lie in its region -- ghost shell included, bit-identical
to the copies other chunks hold -- so each of its halos sees the same sphere as in the whole box."""

import numpy as np

from oracle import mesh as om
from soap_b200 import chunk_tasks as ct
from soap_b200 import synth


def _rows(pos):
    return set(map(bytes, np.ascontiguousarray(pos)))


def test_chunks_hold_the_volumes_particles_of_their_region():
    L = 60.0
    cat = synth.volume_catalogue(300000, 600, L, seed=5, max_np=20000)
    whole, Hw = synth.volume_chunk(cat, cat["index"], full_cover=True)
    pw = whole[1]["Coordinates"].numpy()
    assert len(pw) == 300000 and len(_rows(pw)) == 300000
    H = {k: cat[k] for k in ("cofp", "index", "search_radius", "read_radius", "nr_bound_part", "is_central")}
    hs, cs = ct.peano_decomposition(L, H, 4)
    all_rows = _rows(pw)
    total = 0
    for c in range(4):
        sel = ct.chunk_halos(hs, cs, c)["index"]
        data, halos = synth.volume_chunk(cat, sel)
        pc = data[1]["Coordinates"].numpy()
        total += len(pc)
        assert _rows(pc) <= all_rows  # every particle of the chunk is a particle of the volume, bit for bit
        # the chunk is exactly the volume cut by its cell cover
        keep = ct.ghost_mask(pw, halos["cofp"].numpy(), halos["read_radius"].numpy(), L)
        assert len(pc) == int(keep.sum())
        g = data[1]["GroupNr_bound"].numpy()
        for j in (0, len(sel) // 2, len(sel) - 1):
            h = int(sel[j])
            assert (g == h).sum() == cat["nh"][h] == int(halos["nr_bound_part"][j])
            # same sphere as in the whole volume
            a = om.brute_force_query(pc, halos["cofp"][j].numpy(), float(halos["read_radius"][j]), L)
            b = om.brute_force_query(pw, halos["cofp"][j].numpy(), float(halos["read_radius"][j]), L)
            assert len(a) == len(b) and _rows(pc[a]) == _rows(pw[b])
    assert total > 300000  # ghost shells are duplicated, never exchanged
