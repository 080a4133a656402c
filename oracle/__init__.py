"""TEST INFRASTRUCTURE ONLY: CPU oracle for the GS-SDF hot path (see oracle/splat_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product path (gs-sdf_b200/) never does.
"""
