"""Stand-in for `matplotlib` (TEST INFRASTRUCTURE ONLY)."""
