"""TEST INFRASTRUCTURE ONLY.

CPU restatements (oracles) of the reference algorithms on the 3D_SLN hot path, used by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs as the CHECKER and CPU baseline.  The product (sln_b200/) never
imports anything from here.
"""
