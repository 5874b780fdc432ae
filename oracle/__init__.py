"""CPU oracle for the SELD feature-extraction hot path -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; nothing under seld_b200/ does.
"""
