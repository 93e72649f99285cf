"""Oracle = CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this package (see oracle/reference_path.py for the full statement and the pinning story)."""
