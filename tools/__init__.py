"""Scratch measurement / debugging scripts (not product code).  Run from the repo root: ``python -m tools.<name> [args]``."""
