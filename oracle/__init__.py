"""CPU oracle for the per-frame analyzer hot path.  TEST INFRASTRUCTURE -- see oracle/avd_oracle.c.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
