"""TEST INFRASTRUCTURE ONLY -- CPU/PyTorch restatement of the reference algorithm for the IGGT hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only as the checker.  The product (iggt_official_b200) never imports it.
"""
