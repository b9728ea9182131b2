"""Test-harness stand-in for `codepy` (reference pins codepy>=2019.1,<2025).
NOT PART OF THE PRODUCT — see oracle/refshim/cgen/__init__.py."""
