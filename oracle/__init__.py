"""CPU oracle for the TWA hot path -- test infrastructure only (see oracle/oracle.c)."""
