"""CPU oracle of the BRUNO hot path: TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and the cpu_baseline / reference arm of bench.py; never by the product path)."""
