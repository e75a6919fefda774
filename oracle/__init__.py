"""CPU oracle -- test infrastructure only (see pic_oracle.py header)."""
