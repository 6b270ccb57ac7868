"""TEST INFRASTRUCTURE: CPU oracle for the PBSIM3 hot path (see pbsim_oracle.h)."""
