"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/tc_oracle.c)."""
