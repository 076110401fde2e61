"""
oracle/ -- CPU restatement of SOAP's per-halo particle aggregation hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker (or as the timed
CPU baseline).  Nothing under ``soap_b200/`` imports this package: the product
path runs on the CUDA extension and fails loudly if it is missing.

Parity status: **pinned at the operator level, unpinned at the unit-coercion level.**
The reference's own tests hold no golden numbers for this path and
the reference cannot be imported here (unyt / mpi4py / h5py / virgo are absent), so
``tests/golden/make_golden.py`` compiles the reference's unmodified function bodies
out of /root/reference and executes them under a minimal unyt / MPI stand-in with
all inputs in one unit system; ``tests/test_oracle_golden.py`` checks that this
package reproduces those fixtures bit-exactly (find_SO_radius_and_mass with scipy's
brentq, half-weight radius, Vmax, velocity dispersion, angular momentum,
kappa_corot, cylindrical velocities, 3-D and projected inertia tensors, SharedMesh
arrays and query_radius_periodic sets).  What stays unpinned: the unyt unit
coercions inside the four HaloProperty.calculate bodies and VirgoDC's within-cell
sort order (ports of the reference's invariant tests cover those paths:
mesh brute force ``tests/test_shared_mesh.py:95-125``, half-mass bound
``tests/test_half_mass_radius.py:31``).
Every function cites the reference file:line it
restates.  Units: the reference carries unyt units; here every quantity is a
raw ndarray in *coordinate units* (positions/radii: comoving snap_length,
masses: snap_mass, velocities: snap velocity) and every threshold that unyt
would convert at a comparison is passed in already converted (SURVEY.md 8(c)
detail 11).

Two arithmetic modes are provided where the reference accumulates in float32:
``faithful=True`` restates dtype-for-dtype; ``faithful=False`` accumulates in
float64 (same selections, same operation order otherwise).
"""
