"""Test infrastructure only: CPU oracles for the sparse polynomial multiplier path (see pyoracle.py, piranha_cpu.c)."""
