"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Nothing under ``mjmpc_b200/`` may import this package.  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs (as the checker or the timed CPU baseline).
"""
