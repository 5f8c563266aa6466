"""CPU oracle for the TRAK / Shapley attribution hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement (numpy / CPU torch) of the
reference's arithmetic for the hot path named in BASELINE.json.  It exists to
*check* the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product package (``group-attribution-for-diffusion-models_b200`` a.k.a.
``gadm_b200``) never imports it and fails loudly when ``libgadm.so`` is absent.

Pinning status (see DESIGN.md "Oracle"):

* ``data_shapley`` / ``data_banzhaf`` / ``evaluate_lds`` / scorer restatements are
  pinned against outputs of the reference's own functions imported by file path
  from ``/root/reference`` in the build container; the vectors live in
  ``tests/golden/*.npz`` together with ``tests/golden/make_golden.py``.
* The projector has NO reference-side pin: the reference's projector is the
  third-party ``traker==0.1.3`` + ``fast-jl==0.1.3`` (requirements.txt:10-11),
  absent from /root/reference and not installable here, and the reference has no
  tests or golden vectors.  Projection parity is therefore "parity unpinned"
  w.r.t. trak: it is checked (i) against an explicit-matrix fp64 product using the
  same Philox-generated matrix, (ii) against Philox4x32-10 known-answer vectors
  and (iii) through JL / linearity / tiling-independence properties.
"""
