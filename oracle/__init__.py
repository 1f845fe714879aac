"""CPU oracle of the LoANs STN crop path.  TEST INFRASTRUCTURE ONLY -- see stn_numpy.py / stn_oracle.c.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
