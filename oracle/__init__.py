"""CPU oracle package -- test infrastructure only (see graph_oracle.py header)."""
