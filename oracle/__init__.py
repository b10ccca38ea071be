"""CPU test oracle for the PairHMM path.  TEST INFRASTRUCTURE ONLY (see oracle/pairhmm_oracle.c)."""
