"""Seeded synthetic workloads (C2..C5, divergence / indel sweeps) for bench.py and the tests.
Test and measurement infrastructure -- not part of the product package."""
