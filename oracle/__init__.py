"""CPU oracle package — TEST INFRASTRUCTURE ONLY (see oracle/skm_oracle.h)."""
