"""CPU oracle for the PSMC HMM hot path - test infrastructure, never imported by the product."""
