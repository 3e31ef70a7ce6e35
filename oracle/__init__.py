"""Test infrastructure only -- see oracle/kws_oracle.py."""
