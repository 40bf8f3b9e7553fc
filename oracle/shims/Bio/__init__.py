"""Minimal stand-in for biopython (not installed here) -- test infrastructure only."""
