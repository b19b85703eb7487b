"""TEST INFRASTRUCTURE ONLY: the CPU parity oracle (see mrg_oracle.c, pyoracle.py)."""
