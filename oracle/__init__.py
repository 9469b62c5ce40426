"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's `:exchange` tracer (rthx_oracle.c) plus the numpy consumers needed to pin it
against the reference's golden vectors (grey_solver.py).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` leg may import this package; the product path never does.
"""
