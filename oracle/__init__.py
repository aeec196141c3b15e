"""TEST INFRASTRUCTURE ONLY - the parity checker (C restatement + reference harness)."""
