"""CPU checkers for the inflate / PNG / gzip hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package. The product (debigulator_b200) never
does; it fails loudly when its CUDA library is missing.
"""
