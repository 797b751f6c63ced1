"""TEST INFRASTRUCTURE: CPU oracle for the MoRig rigging-network forward path.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this package, and only as the checker / reported CPU baseline.
`morig_b200` never imports it (enforced by tests/test_boundary.py).
"""
