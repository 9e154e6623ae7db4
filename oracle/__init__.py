"""Test infrastructure only: CPU oracle for the JaggedArray hot path (see jagged_oracle.py)."""
