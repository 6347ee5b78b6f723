"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the RoDynRF ray-march hot path (see rodynrf_oracle.py).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product path (robust-dynrf_b200/) never does.
"""
