"""oracle/ -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).

CPU restatement of the reference's hot path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` leg may import this package.  The product
(avatarcraft_b200/) never does: it fails loudly when its CUDA library is missing.
"""
