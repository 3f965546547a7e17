"""CPU oracle for the hot path -- TEST INFRASTRUCTURE ONLY.

`pyref` is the big-integer cross-oracle; `coracle` wraps the C restatement (oracle/zkp_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
