"""Placeholder of the reference's (empty) simulation-runner module (pyapes/runner.py)."""
