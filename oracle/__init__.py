"""oracle/ — CPU restatement + runner of the reference's gradient-sync path.

TEST INFRASTRUCTURE ONLY.  Nothing under ray_lightning_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
"""
