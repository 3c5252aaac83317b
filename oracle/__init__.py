"""TEST INFRASTRUCTURE ONLY: CPU oracle for the hot path (see bb_oracle.c / pyref.py headers)."""
