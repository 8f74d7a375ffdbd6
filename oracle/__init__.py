"""ORACLE — TEST INFRASTRUCTURE ONLY. See oracle/yk_oracle.h."""
