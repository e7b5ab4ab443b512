"""CPU oracle for the image-realism hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or the timed CPU baseline.  The product package (``tise_toolbox_b200``) never imports
it and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests and no golden vectors for this
path, so the oracle is pinned by (1) ``oracle/reference_shim.py`` -- the reference's own
``fid_score.py`` / ``inception.py`` / O-IS source imported unmodified from ``/root/reference``
(possible only in the build container) and compared against ``oracle/restate.py`` on the same
inputs by ``tools/make_golden.py``, which also writes the committed fixtures under
``tests/golden/``; (2) the one transferable known-answer of the reference's result files,
``FID(X, X) ~ 0`` (``object_fidelity/results/O-FID/real_images.txt``).
For the rows added around the path the checker is the routine the reference itself calls: Pillow's
PNG / JPEG decoders and ``Image.crop`` (bit-exact comparisons in ``tests/``), and the reference's own
``temperature_scaling.py`` / ``ranking_score.py`` run by ``tools/make_golden_calibration.py`` into
``tests/golden/calibration.json``.
"""
