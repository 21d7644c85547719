"""CPU oracle for the MODWT/SWT hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (vectorwave_b200/) never does.

Parity status: pinned by the reference's own known-answer tests and formulas
(see modwt_oracle.c header and tests/test_oracle_kat.py).

Contents
  modwt_oracle.c   C restatement of the reference loops (file:line cited per function)
  cref.py          ctypes loader + thin numpy-facing wrappers for the C oracle
  nptwin.py        independent numpy restatement (same roundings) used to cross-check the C
  wavelets.py      filter tables transcribed as data
  javarandom.py    java.util.Random restatement for regenerating the reference's fixtures
"""
