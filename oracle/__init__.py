"""CPU oracle (test infrastructure only; see dsin_oracle.py header)."""
