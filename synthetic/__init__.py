"""Seeded synthetic pages and weights shared by the tests, the oracle and bench.py."""
