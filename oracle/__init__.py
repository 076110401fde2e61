"""
oracle/ -- CPU restatement of SOAP's per-halo particle aggregation hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker (or as the timed
CPU baseline).  Nothing under ``soap_b200/`` imports this package: the product
path runs on the CUDA extension and fails loudly if it is missing.

Parity status: **parity unpinned by stored golden numbers** -- the reference's
own tests hold no golden vectors for this path, and the
reference cannot be imported here (unyt / mpi4py / h5py / virgo are absent).
The oracle is therefore pinned by (1) ports of the reference's invariant tests
(mesh brute force ``tests/test_shared_mesh.py:95-125``, half-mass bound
``tests/test_half_mass_radius.py:31``, NFW concentration within 10 %
``tests/test_SO_properties.py:434-446``) and (2) direct use of the same two
third-party numeric routines the reference calls (``scipy.optimize.brentq``,
``numpy.linalg.eigh``).  Every function cites the reference file:line it
restates.  Units: the reference carries unyt units; here every quantity is a
raw ndarray in *coordinate units* (positions/radii: comoving snap_length,
masses: snap_mass, velocities: snap velocity) and every threshold that unyt
would convert at a comparison is passed in already converted (SURVEY.md 8(c)
detail 11).

Two arithmetic modes are provided where the reference accumulates in float32:
``faithful=True`` restates dtype-for-dtype; ``faithful=False`` accumulates in
float64 (same selections, same operation order otherwise).
"""
