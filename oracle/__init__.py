"""TEST INFRASTRUCTURE — CPU oracle of the LTE hot path (see oracle/lte_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. The product (geodesicodis_b200) never does.
"""
