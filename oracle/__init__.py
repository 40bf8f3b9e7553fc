"""CPU oracle -- test infrastructure only (see oracle/asref.c header)."""
