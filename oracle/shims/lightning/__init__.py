"""ORACLE / TEST INFRASTRUCTURE ONLY — minimal stand-in for `lightning` so the reference's duett/duett.py imports."""
