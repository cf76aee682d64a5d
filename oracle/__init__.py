"""CPU oracle for the GEOSmie Mie lookup-table hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product (geosmie_b200/) never does.
"""
