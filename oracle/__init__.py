"""Test-infrastructure oracles (CPU restatements of the reference algorithm).  NOT product code."""
