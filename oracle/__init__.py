"""oracle/ -- CPU restatement of the reference's likelihood hot path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product
package `varycoef_b200` never does (tests/test_no_oracle_in_product.py enforces it).
"""
