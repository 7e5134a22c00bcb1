"""CPU oracles (test infrastructure).  Never imported by bfsr_b200/."""
