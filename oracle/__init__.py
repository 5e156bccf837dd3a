"""ORACLE — test infrastructure only.  Nothing under esr_nerf_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do."""
