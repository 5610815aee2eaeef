"""ORACLE - TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's hot path used as the
parity checker.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product
(smap_b200/) never does.
"""
