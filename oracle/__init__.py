"""TEST INFRASTRUCTURE — CPU oracle for the NexToU hot path (see oracle/README.md).

Nothing under nextou_b200/ may import this package; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do.
"""
