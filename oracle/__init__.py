"""CPU oracle for the gsr hot path -- TEST INFRASTRUCTURE ONLY (see gsr_oracle.c header)."""
