"""ORACLE package: CPU restatements of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under wavenet_autoencoders_b200/ imports this; only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py do (as the checker / the thing timed on host cores).
"""
