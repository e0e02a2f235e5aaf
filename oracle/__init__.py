"""CPU oracle for the hot path (test infrastructure).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import anything from this package.  The product path (``gnn_motion_planning_b200``) never
does, and has no CPU fallback.
"""
