"""CPU oracle (test infrastructure only — see pa_oracle.py header)."""
