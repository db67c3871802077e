"""CPU oracle for the per-RoI captioning hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``image-captioning_b200/`` imports this package.  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py``.  PARITY UNPINNED (the reference has no tests and cannot be run here); see the
module headers.
"""
