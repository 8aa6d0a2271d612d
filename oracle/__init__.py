"""CPU oracle for the StringWars similarities hot path — TEST INFRASTRUCTURE ONLY (see lev_oracle.c)."""
