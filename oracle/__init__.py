"""CPU oracle (test infrastructure only) -- see oracle/qstep_oracle.c."""
