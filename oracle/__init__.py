"""CPU oracle (test infrastructure only; see retrieval_oracle.py and radar_oracle.c)."""
