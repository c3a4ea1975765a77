"""Parity oracles for branson_b200 -- TEST INFRASTRUCTURE ONLY.

Nothing under this package is product code.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.
"""
