"""CPU oracle for flexs_b200 — TEST INFRASTRUCTURE ONLY (see oracle/flexs_oracle.py header).

Nothing under ``flexs_b200/`` imports this package.
"""
