"""CPU checkers for the render path. TEST INFRASTRUCTURE ONLY -- see oracle/rt_oracle.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package.
"""
