"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's encode/decode hot path, used exclusively as the *checker*:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing in l3c_pytorch_b200/ imports it (tests/test_no_oracle_in_product.py
greps for that).

Pinning status: the coder + CDF restatement (ac_oracle.c) is pinned byte-for-byte against the
reference's own torchac.cpp compiled unmodified (oracle/_ref, see pin_oracle.py) and against the
known-answer vectors of SURVEY.md section 8c; the network/DMLL restatement (model.py) is pinned
against the unmodified reference Python imported from /root/reference (gen_golden.py writes
tests/golden/*).  Pillow's bicubic (RGB baselines) is third-party: parity unpinned by any
reference test.
"""
