"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's arithmetic for the GP hot path.  Nothing under ``treegp_b200/``
may import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker / the timed CPU baseline.
"""
