"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the aesmc SMC hot path (see smc_oracle.c and reference_port.py headers).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; nothing under aesmc_b200/ does.
"""
