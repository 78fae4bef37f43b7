"""CPU oracle for the clip-transform hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the timed CPU arm.
The product package ``vision_collision_detection_b200`` never imports it and
fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the unmodified
reference itself, imported in the build container by ``oracle/ref_import.py``
and frozen by ``tests/golden/make_golden.py`` into ``tests/golden/*.npz``
(torch 2.11.0 / torchvision 0.26.0).  The sampler / sensor restatements are pinned the
same way against the unmodified reference Datasets (``tests/golden/make_dataset_golden.py``
-> ``dataset_golden.json``, ``dataset_sensor.npz``).
"""
