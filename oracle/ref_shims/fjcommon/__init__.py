"""Stand-in for the reference's un-vendored dependency fjcommon==0.2.10 (pip_requirements.txt:4).

TEST INFRASTRUCTURE ONLY: lets /root/reference/src import in this container so that golden
vectors can be generated from the unmodified reference (oracle/gen_golden.py).  Only the handful
of helpers the encode/decode path touches are provided."""
