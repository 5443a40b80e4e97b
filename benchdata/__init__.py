"""Synthetic "msprime-shaped" benchmark inputs.  Shared by bench.py,
__graft_entry__.smoke() and the tests; contains no reference code and no oracle code."""
