"""TEST INFRASTRUCTURE ONLY -- CPU restatements of the reference's two hot paths.

Nothing under graphrole_b200/ imports this package.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs may use it, and only as the checker or the
CPU baseline, never as the thing shipped.
"""
