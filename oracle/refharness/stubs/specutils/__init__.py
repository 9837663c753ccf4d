"""Stand-in for `specutils` (TEST INFRASTRUCTURE ONLY; import-time use in the reference's utils.py:9)."""
