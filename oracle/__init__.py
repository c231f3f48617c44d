"""CPU oracle — test infrastructure, never imported by lattice_qft_b200 (see lqft_oracle.c)."""
