"""CPU oracle (test infrastructure only; see unimatch_oracle.py)."""
