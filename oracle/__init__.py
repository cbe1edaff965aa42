"""CPU oracle for the FOCAL contrastive-loss hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline.  ``focal_b200`` never imports this package.
"""
