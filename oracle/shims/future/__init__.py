"""Stand-in for the `future` package (test infrastructure only)."""
