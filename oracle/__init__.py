"""CPU oracle package -- test infrastructure only (see oracle/vq_oracle.py header)."""
