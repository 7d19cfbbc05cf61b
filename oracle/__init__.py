"""TEST INFRASTRUCTURE — CPU oracle of the LM inner step (see oracle/ba_oracle.hpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
