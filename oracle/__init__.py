"""CPU oracle for the CFR hot path -- TEST INFRASTRUCTURE ONLY (see cfr_oracle.c header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
