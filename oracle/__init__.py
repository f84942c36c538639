"""CPU oracle for the COVINS hot path — TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  covins_b200/ must never import this package (tests/test_boundary.py enforces it).
"""
